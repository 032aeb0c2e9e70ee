"""Drop-in modules for the reference's reasoning path, backed by libdfol_b200.

  FastBoxFeaturizer     <-> BatchGQABoxFeaturizer   (reference: src/nsvqa/data/batch_gqa_boxfeatures_pipeline.py:193-281)
  FastClassifierOracle  <-> ClassifierOracle        (reference: src/nsvqa/nn/vision/classifier_oracle.py:11-156)
  FastGQAInterpreter    <-> BatchGQAInterpreter     (reference: src/nsvqa/nn/interpreter/batch_gqa_interpreter.py:13-86,
                                                     batch_base_interpreter.py:14-183)

Same constructor arguments, same ``forward(program_batch_list, is_training, return_trace, modulator_switch)`` and
result dict, same state-dict key names (the networks are held under the same attribute paths), so they can be
swapped in under gqa_interpreter_experiments.py.  ``result['log_probability']`` is attached to
the autograd graph through one custom Function whose backward runs the hand-written backward kernels, so the
reference trainer's loss / ``loss.backward()`` / clip / Adam work unchanged.  ``FusedTrainStep`` is the fast
path: loss, gradient all-reduce, clip and Adam without leaving our kernels.
"""

import math
import os

import numpy as np
import torch
import torch.nn as nn

from . import capi
from .capi import K, call, ptr
from .compiler import BINARY, QUERY, STATEMENT, ProgramCompiler
from .engine import OracleWeights, ReasoningEngine, SceneLayout
from .modulator_cuda import NativeAttentionTransfer
from .networks import dropout_p, linear_layers
from .parallel import FlatBucket

YES = ('yes', 'yeah', 'yep', 'yup', 'aye', 'yea')


class FastBoxFeaturizer(nn.Module):

    def __init__(self, featurizer_network=None):
        super(FastBoxFeaturizer, self).__init__()
        if featurizer_network is None:
            raise NotImplementedError('the fused path needs the featurizer network (Linear + Sigmoid)')
        self._featurizer_network = featurizer_network

    def featurize_scene(self, device, objects_list, batch_index, meta_data=None):
        """BatchGQABoxFeaturizer.featurize_scene (reference nsvqa/data/batch_gqa_boxfeatures_pipeline.py:199-281), default
        branch (no pre-featurised relations, no ``object_pairs`` supervision), on the CUDA kernels -- forward only.

        Returns the reference's dict: ``attribute_features`` (T, F + 4) = [sigmoid(X Wf^T + b) | box position],
        ``relation_features`` = {'features': (P', 2 (F + 4) + 4) rows [obj_s | obj_o | dist | asin | sign x | sign y] of
        all ordered pairs (s, o), s != o, of each image, 'index': [image, s, o] (global object rows)} in the order of
        util.find_sparse_pair_indices (:87-103: subject-major), and ``object_num``.  FastGQAInterpreter.forward never
        materialises this matrix (the first relation layer is evaluated as U[s] + V[o] + Wg.geo); this method exists for
        callers that use the featurizer on its own, as BatchInterpreterBase.build_scene does."""
        from .capi import K, call, ptr, stream_ptr
        from .engine import SceneLayout, gemm_f32
        from .networks import linear_layers
        meta_data = meta_data or {}
        if 'relation_features' in meta_data or 'object_pairs' in meta_data:
            raise NotImplementedError('pre-featurised relations / object_pairs supervision are outside the fused path')
        (feat,) = linear_layers(self._featurizer_network)
        x = objects_list.to(device=device, dtype=torch.float32).contiguous()
        assert x.is_cuda, 'dfol_vqa_b200 runs on CUDA tensors only (no CPU fallback)'
        T, D = x.shape[0], x.shape[1] - 6
        F = feat.weight.shape[0]
        ldo = F + 4
        st = stream_ptr(x.device)
        obj = torch.empty(T, ldo, device=x.device, dtype=torch.float32)
        with torch.no_grad():
            gemm_f32(x[:, :D], feat.weight.t(), obj[:, :F], feat.bias, K.ACT_SIGMOID, stream=st)
            call('dfol_box_position', ptr(x), x.stride(0), D, ptr(obj), ldo, F, T, st)
            counts = torch.bincount(batch_index.to(x.device)).tolist()
            lay = SceneLayout.get(counts, 1, 1, x.device)
            width = 2 * ldo + 4
            pm = torch.empty(lay.P, width, device=x.device, dtype=torch.float32)
            if lay.P:
                call('dfol_pair_features_dropout', ptr(obj), ldo, ldo, F, ptr(pm), width, width, 0, ptr(lay.pair_row),
                     ptr(lay.obj_row), ptr(lay.img_n), ptr(lay.pair_img), lay.P, 0, 0, 0.0, st)
            # (s, o) of every pair row, self pairs dropped: index bookkeeping of the reference's return value
            img = lay.pair_img.long()
            n = lay.img_n.long()[img]
            local = torch.arange(lay.P, device=x.device) - lay.pair_row.long()[img]
            s_loc, o_loc = local // n.clamp(min=1), local % n.clamp(min=1)
            keep = s_loc != o_loc
            base = lay.obj_row.long()[img]
            index = [img[keep], (base + s_loc)[keep], (base + o_loc)[keep]]
            relation = pm[keep] if bool(keep.any()) else None
        return {'attribute_features': obj, 'relation_features': {'features': relation, 'index': index}, 'object_num': T}


class FastClassifierOracle(nn.Module):

    def __init__(self, ontology, attribute_network, relation_network, embedding_network, normalize=False, cached=False):
        super(FastClassifierOracle, self).__init__()
        self._ontology = ontology
        self._feature_dim = 1
        self._attribute_network = attribute_network
        self._relation_network = relation_network
        self._embedding_network = embedding_network
        self._normalize = normalize
        self._cached = cached

    def get_embedding(self, tokens, meta_data, device):
        """OracleBase.get_embedding (reference nsvqa/nn/vision/base_oracle.py:45-55): word embeddings of concept tokens
        (the calibrator's input), from the batch's pre-gathered table when it has one, else from the ontology."""
        import numpy as np
        if meta_data is not None:
            try:
                ind = [meta_data['index'][t] for t in tokens]
                return meta_data['embedding'][ind, :]
            except KeyError:
                pass
        return torch.from_numpy(np.asarray(self._ontology.get_embeddings(tokens))).float().to(device)

    def compute_all_log_likelihood_2(self, object_features, pair_object_features):
        """ClassifierOracle.compute_all_log_likelihood_2 (reference nsvqa/nn/vision/classifier_oracle.py:145-156) on the
        CUDA GEMM kernels, forward only: (T, C) log-likelihoods of every concept for every object and (P', R)
        log-likelihoods of every relation for every pair row (the embedding layer restricted to
        ``ontology._relation_index``, i.e. the columns the reference slices).  Dense row-major results, as the reference
        returns them; FastGQAInterpreter.forward uses the per-image table layout and demand-driven relation columns
        instead."""
        from .capi import K, stream_ptr
        from .engine import gemm_f32
        from .networks import linear_layers
        (emb,) = linear_layers(self._embedding_network)

        def chain(net, h, weight, bias):
            layers = linear_layers(net)
            st = stream_ptr(h.device)
            for i, layer in enumerate(layers):
                out = torch.empty(h.shape[0], layer.weight.shape[0], device=h.device, dtype=torch.float32)
                gemm_f32(h, layer.weight.t(), out, layer.bias, K.ACT_ELU if i < len(layers) - 1 else K.ACT_SIGMOID,
                         stream=st)
                h = out
            out = torch.empty(h.shape[0], weight.shape[0], device=h.device, dtype=torch.float32)
            gemm_f32(h, weight.t(), out, bias, K.ACT_LOGSIGMOID, stream=st)
            return out

        with torch.no_grad():
            x = object_features.float().contiguous()
            assert x.is_cuda, 'dfol_vqa_b200 runs on CUDA tensors only (no CPU fallback)'
            attr = chain(self._attribute_network, x, emb.weight, emb.bias)
            rel = None
            if pair_object_features is not None:
                ridx = torch.as_tensor(self._ontology._relation_index, device=x.device, dtype=torch.long)
                rel = chain(self._relation_network, pair_object_features.float().contiguous(),
                            emb.weight[ridx].contiguous(), emb.bias[ridx].contiguous())
        return attr, rel


class _ReasoningFunction(torch.autograd.Function):
    """features + oracle parameters (+ attention-network parameters) -> log-probabilities of one program batch."""

    @staticmethod
    def forward(ctx, engine, cp, layout, features, need_grad, attention, sink, dropout, n_oracle, *params):
        oracle_params = params[:n_oracle]
        oracle_grad = need_grad and any(p.requires_grad for p in oracle_params)
        scene = engine.build_scene(features, layout, keep_for_backward=oracle_grad, cp=cp, dropout=dropout)
        mod_ctx = None
        if attention is not None:
            # token-side LSTM passes of the calibrator (libdfol_b200 kernels, modulator_cuda.py) -> one row of
            # (alpha, beta, c, d) per predicate, consumed by the interpreter kernels
            scene.mods, mod_ctx = attention.forward(cp)
        lp, tape = engine.run_programs(cp, scene, save_tape=need_grad or sink is not None)
        if sink is not None:
            sink['tape'] = tape
        ctx.engine, ctx.cp, ctx.scene, ctx.tape, ctx.params = engine, cp, scene, tape, params
        ctx.oracle_grad, ctx.n_oracle, ctx.attention, ctx.mod_ctx = oracle_grad, n_oracle, attention, mod_ctx
        return lp

    @staticmethod
    def backward(ctx, d_lp):
        params, scene, n_oracle = ctx.params, ctx.scene, ctx.n_oracle
        d_lp = d_lp.contiguous().float()
        if ctx.mod_ctx is not None:
            scene.d_mods = torch.zeros_like(scene.mods)
        grads = {id(p): torch.zeros_like(p, dtype=torch.float32)
                 for k, p in enumerate(params) if (ctx.oracle_grad if k < n_oracle else ctx.mod_ctx is not None)}
        if ctx.oracle_grad:
            ctx.engine.backward(ctx.cp, scene, ctx.tape, d_lp, grads)
        else:
            # frozen oracle (sample_config.yaml: only the attention networks train): the backward interpreter alone
            # yields d loss / d modulations; the scene's backward pass is skipped
            ctx.engine.program_backward(ctx.cp, scene, ctx.tape, d_lp)
        if ctx.mod_ctx is not None and any(p.requires_grad for p in params[n_oracle:]):
            ctx.attention.backward(ctx.mod_ctx, scene.d_mods, grads)
        ctx.scene = ctx.tape = ctx.mod_ctx = None
        out = []
        for k, p in enumerate(params):
            live = p.requires_grad and (ctx.oracle_grad if k < n_oracle else ctx.attention is not None)
            out.append(grads[id(p)] if live else None)
        return (None, None, None, None, None, None, None, None, None) + tuple(out)


class _TraceEntry(object):
    """What the reference's trace consumers read of a BatchVariableSet (batch_base_types.py:34-100)."""

    def __init__(self, log_attention, names):
        self._log_attention = log_attention
        self._name = names

    def get_attention(self):
        return self._log_attention.exp()


class _Holder(nn.Module):
    """Registers shared sub-modules under the reference's attribute path (state-dict key compatibility)."""

    def __init__(self, **children):
        super(_Holder, self).__init__()
        for k, v in children.items():
            setattr(self, k, v)


class FastGQAInterpreter(nn.Module):

    def __init__(self, name, oracle, ontology, featurizer=None, trainable_module_type=None, feature_dim=1,
                 trainable_gate=False, likelihood_threshold=0, hard_mode=False, attention_transfer_state_dim=0,
                 forward_attention_network=None, backward_attention_network=None, attention_output_network=None,
                 apply_modulation_everywhere=True, cached=False, visual_rule_learner=None, calibrator=None,
                 gemm_mode=None):
        super(FastGQAInterpreter, self).__init__()
        if trainable_module_type is not None or trainable_gate or feature_dim != 1:
            raise NotImplementedError('trainable logic gates / operator MLPs are outside the hot path (SURVEY.md §2)')
        attention_nets = (forward_attention_network, backward_attention_network, attention_output_network)
        if any(n is not None for n in attention_nets) and not all(n is not None for n in attention_nets):
            raise ValueError('the attention-transfer calibrator needs all three attention networks')
        if not apply_modulation_everywhere and attention_nets[0] is not None:
            raise NotImplementedError('apply_modulation_everywhere=False (the reference itself fails on it: '
                                      'batch_base_interpreter.py:90-92 sets an attribute on a list)')
        if visual_rule_learner is not None or calibrator is not None:
            raise NotImplementedError('visual rule learner / calibrator are not part of the reference hot path')
        if featurizer is None:
            raise NotImplementedError('the fused path needs the box featurizer')
        self._name = name
        self._featurizer = featurizer
        self._oracle = oracle
        self._ontology = ontology
        self._likelihood_threshold = likelihood_threshold
        self._hard_mode = hard_mode
        self._global_step = nn.Parameter(torch.tensor([0], dtype=torch.float), requires_grad=False)
        self._has_modulator = attention_nets[0] is not None
        self._attention_transfer_state_dim = attention_transfer_state_dim
        self._cached = cached
        self._gemm_mode = gemm_mode or os.environ.get('DFOL_GEMM_MODE', 'fp32')

        for net in (featurizer._featurizer_network, oracle._attribute_network, oracle._relation_network,
                    oracle._embedding_network):
            if dropout_p(net) > 0:
                self._dropout = dropout_p(net)
        self._weights = OracleWeights(linear_layers(featurizer._featurizer_network),
                                      linear_layers(oracle._attribute_network),
                                      linear_layers(oracle._relation_network),
                                      linear_layers(oracle._embedding_network)[0])
        self._engine = ReasoningEngine(self._weights, ontology._relation_index, self._gemm_mode)
        self._compiler = ProgramCompiler(ontology, normalize=oracle._normalize, hard_mode=hard_mode,
                                         relation_slots=(self._gemm_mode == 'bf16'), modulated=self._has_modulator,
                                         concept_num=self._weights.emb.weight.shape[0])
        self._attention = None
        if self._has_modulator:
            # the reference registers the SAME three networks under every operator module; the first registration
            # (_ops.select._filter.*) names the parameters, the remaining paths are kept so that checkpoints written by
            # either side load into the other (load_state_dict(strict=False), batch_base_interpreter.py:42-43)
            def flt():
                return _Holder(_forward_attention_network=forward_attention_network,
                               _backward_attention_network=backward_attention_network,
                               _attention_output_network=attention_output_network)
            def sel():
                return _Holder(_filter=flt())
            def rel():
                return _Holder(_gqa_select=sel(), _relate=flt())
            self._ops = nn.ModuleDict({
                'select': sel(), 'filter': sel(), 'relate': rel(), 'query_attr': _Holder(_gqa_choose_attr=sel()),
                'choose_attr': sel(), 'verify_attrs': sel(), 'choose_rel': rel(),
                'verify_rel': _Holder(_gqa_relate=rel()), 'all_same': sel(),
                'all_different': _Holder(_gqa_all_same=sel()), 'two_same': sel(),
                'two_different': _Holder(_gqa_two_same=sel()), 'compare': sel()})
            self._attention = NativeAttentionTransfer(forward_attention_network, backward_attention_network,
                                                      attention_output_network, ontology)

    _dropout = 0.0

    # ---- reference surface -------------------------------------------------------------------------------

    def parameter_count(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def save(self, export_path_base):
        torch.save(self.state_dict(), os.path.join(export_path_base, self._name))

    def load(self, import_path_base):
        self.load_state_dict(torch.load(os.path.join(import_path_base, self._name)), strict=False)

    def oracle_parameters(self):
        return self._weights.parameters()

    def attention_parameters(self):
        return [] if self._attention is None else self._attention.parameters()

    def modulator(self, cp, modulator_switch=True):
        """The calibrator to run for a compiled batch, or None (no calibrator / switched off / nothing to modulate)."""
        if self._attention is None or not modulator_switch or cp.mod_rows == 0:
            return None
        return self._attention

    def modulations(self, cp, modulator_switch=True):
        """(cp.mod_rows, 4) modulations of a compiled batch (no autograd), or None."""
        att = self.modulator(cp, modulator_switch)
        return None if att is None else att.forward(cp)[0]

    # ---- helpers -----------------------------------------------------------------------------------------

    @staticmethod
    def _object_counts(pb):
        cached = getattr(pb, '_dfol_counts', None)
        if cached is not None:
            return cached
        bidx = pb._object_batch_index
        host = bidx.cpu() if bidx.is_cuda else bidx
        host = host.to(torch.int64)
        assert bool((host[1:] >= host[:-1]).all()), 'object rows must be grouped by image'
        counts = torch.bincount(host).tolist()
        pb._dfol_counts = counts
        src = getattr(pb, '_dfol_host', None)
        if src is not None:
            src._dfol_counts = counts
        return counts

    def compiled(self, pb, give_answer):
        cache = getattr(pb, '_dfol_compiled', None)
        if cache is None:
            cache = {}
            pb._dfol_compiled = cache
            host = getattr(pb, '_dfol_host', None)
            if host is not None:
                host._dfol_compiled = cache
        # (the bytecode depends on the compiler's table layout and on whether modulation rows are assigned)
        key = (bool(give_answer and self._hard_mode), self._compiler.relation_slots, self._compiler.modulated,
               self._compiler.demand_pairs)
        if key not in cache:
            cache[key] = self._compiler.compile(pb, self._object_counts(pb), give_answer=give_answer)
        return cache[key]

    def _features_of(self, pb):
        """The box features of a program batch on the device: the reference's (T, D+6) fp32 tensor, or -- when the batch
        was staged with ProgramBatch.stage_bf16 -- the pair (bf16 features (T, D), fp32 geometry (T, 6))."""
        staged = getattr(pb, '_staged', None)
        if staged is not None and pb._object_features is None:
            if self._gemm_mode != 'bf16':
                raise RuntimeError('bf16-staged program batches need the tensor-core mode (gemm_mode="bf16")')
            if not staged[0].is_cuda:
                raise RuntimeError('dfol_vqa_b200 runs on CUDA only (no CPU fallback): move the program batch to the '
                                   'GPU first (ProgramBatch.to_cuda)')
            return staged
        feats = pb._object_features
        if not feats.is_cuda:
            raise RuntimeError('dfol_vqa_b200 runs on CUDA only (no CPU fallback): move the program batch to the '
                               'GPU first (ProgramBatch.to_cuda)')
        return feats.float().contiguous()

    def build_scene(self, device, object_features, batch_index, meta_data=None, is_training=False):
        """BatchInterpreterBase.build_scene (reference nsvqa/nn/interpreter/batch_base_interpreter.py:45-70): featurizer +
        visual oracle of a collated batch of images.  Returns the engine's ``Scene`` (per-image attribute / relation
        log-likelihood tables in the layout of include/dfol_b200.h: the role of the reference's BatchWorld) instead of
        dense (T, C) / (P', R) tensors.  Without a compiled program batch every relation column is evaluated
        (fp32 parity mode; the tensor-core mode builds its demand-driven relation slots from the programs, in
        ``forward``)."""
        from .engine import SceneLayout
        feats = object_features.to(device=device, dtype=torch.float32).contiguous()
        assert feats.is_cuda, 'dfol_vqa_b200 runs on CUDA tensors only (no CPU fallback)'
        counts = torch.bincount(batch_index.to(feats.device)).tolist()
        w = self._weights
        lay = SceneLayout.get(counts, w.emb.weight.shape[0], len(self._engine.rel_index_host), feats.device)
        with torch.no_grad():
            return self._engine.build_scene(feats, lay, keep_for_backward=is_training)

    def _dropout_for(self, is_training):
        """None, or (p, seed) of this forward pass: nn.Dropout is active when the module is in train() mode and the
        networks were built with dropout > 0 (sample_config.yaml: 0.1).  With frozen oracle networks (the sample
        configuration: only the attention networks train) only the forward pass is masked; with trainable ones both
        modes also run the backward pass of the masked layers.  A fresh seed per call is drawn from torch's CPU generator
        (reproducible under torch.manual_seed).  The reference's own RNG stream cannot be matched;
        the masks are a pure function of the seed (csrc/dropout_kernels.cu) and the tests export them for the oracle."""
        if not (is_training and self.training and self._dropout > 0):
            return None
        seed = self._fixed_dropout_seed
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        self._last_dropout_seed = seed
        return (float(self._dropout), seed)

    _fixed_dropout_seed = None
    _last_dropout_seed = None

    # ---- forward -----------------------------------------------------------------------------------------

    def forward(self, program_batch_list, is_training, return_trace=False, modulator_switch=True):
        dropout = self._dropout_for(is_training)
        give_answer = not is_training
        params = self._weights.parameters()
        need_grad = is_training and torch.is_grad_enabled() and any(
            p.requires_grad for p in params + (self.attention_parameters() if modulator_switch else []))
        lps, metas, traces = [], [], []
        for k, pb in enumerate(program_batch_list):
            drop_k = None if dropout is None else (dropout[0], dropout[1] + k)  # independent masks per sub-batch
            feats = self._features_of(pb)
            cp = self.compiled(pb, give_answer)
            layout = SceneLayout.of_compiled(cp, feats[1].device if isinstance(feats, tuple) else feats.device,
                                             dense_pairs=drop_k is not None)
            att = self.modulator(cp, modulator_switch)
            sink = {} if return_trace else None
            lp = _ReasoningFunction.apply(self._engine, cp, layout, feats, need_grad, att, sink, drop_k, len(params),
                                          *(params + (self.attention_parameters() if att is not None else [])))
            lps.append(lp)
            metas.append(cp)
            traces.append(self._trace(cp, layout, sink['tape']) if return_trace else [])
        result = self._gather(lps, metas, give_answer)
        if return_trace:
            return result, traces
        return result

    def _trace(self, cp, layout, tape):
        """Per-slot attention states for ``return_trace=True`` (VQATrainer._visualize_batch reads
        ``trace[k][i]._log_attention[j, :]``, trainer.py:548, :591): entry i holds the (questions, T) log-attention after
        op slot i, rebuilt from the interpreter's attention tape.  Objects of other images -- which the reference fills
        with by-products that no output depends on -- are set to log(1e-20).  The terminal slot has no entry."""
        B, T, dev = layout.B, layout.T, tape.device
        stride = tape.numel() // max(cp.instr.shape[0], 1)
        rows = tape.view(-1, stride)
        q_first = torch.from_numpy(cp.q_instr[:-1].astype(np.int64)).to(dev)
        n = layout.img_n.long()
        col = torch.arange(stride, device=dev)[None, :].expand(B, stride)
        valid = col < n[:, None]
        dst = (layout.obj_row[:-1].long()[:, None] + col)[valid]
        qi = torch.arange(B, device=dev)[:, None].expand(B, stride)[valid]
        out = []
        for i in range(cp.slot_after.shape[0]):
            ip = q_first + torch.from_numpy(cp.slot_after[i]).to(dev)
            att = torch.full((B, T), math.log(1e-20), device=dev, dtype=torch.float32)
            att[qi, dst] = rows[ip][valid]
            out.append(_TraceEntry(att, cp.slot_names[i]))
        return out

    def _gather(self, lps, metas, give_answer):
        """gather_results (reference: nsvqa/nn/interpreter/data_parallel.py:15-50) + host-side answers."""
        kind = metas[0].kind
        lp = lps[0] if len(lps) == 1 else torch.cat(lps)
        answer, answer_lp, options = [], [], []
        if kind == QUERY:
            for cp in metas:
                options += [list(o) if not isinstance(o, tuple) else o for o in cp.options]
        elif kind == BINARY:
            options = ['no', 'yes']
        if give_answer:
            start = 0
            for cp in metas:
                a, alp = self._answers_device(cp, lp.detach()[start:start + cp.lp_num])
                answer += a
                answer_lp += alp
                start += cp.lp_num
        return {'answer': answer, 'log_probability': lp, 'options': options, 'variable_set': None, 'type': kind,
                'cumulative_loss': 0, 'variable_sets_num': 0, 'answer_log_probability': answer_lp}

    def _answers_device(self, cp, lp):
        """Answer lists of one compiled batch with the decision made ON THE DEVICE (dfol_answers: yes / no, argmax with
        find_max_ind's exact-tie semantics and the likelihood threshold, compare): the host reads 12 bytes per question
        and builds the answer strings; the membership bytes and log-probabilities of a question are only fetched when
        its answer is not a single option (exact ties).  Reference: batch_gqa_ops.py:222-225, :404-407, :744-748,
        util.py:64-66."""
        if cp.kind == STATEMENT:
            return [[n] for n in cp.names], []
        dev = lp.device
        Q = cp.question_num
        mode = 0 if cp.kind == BINARY else (2 if cp.terminal == 'compare' else 1)
        seg = None if mode == 0 else self._engine.upload_programs(cp, dev)['seg']
        out = torch.empty(3 * Q, device=dev, dtype=torch.int32)
        sel = torch.empty(cp.lp_num, device=dev, dtype=torch.uint8) if mode == 1 else None
        lp = lp.contiguous()
        call('dfol_answers', ptr(lp), ptr(seg), Q, mode, float(self._likelihood_threshold), ptr(out[:Q]),
             ptr(out[Q:2 * Q]), ptr(out[2 * Q:]), ptr(sel), capi.stream_ptr(dev))
        host = out.cpu().numpy()
        first, count = host[:Q], host[Q:2 * Q]
        best = host[2 * Q:].view(np.float32)
        if mode == 0:
            p = np.exp(best)
            ans = [['yes'] if f else ['no'] for f in first]
            alp = [[math.log(float(v))] if f else [math.log(1.0 - float(v))] for v, f in zip(p, first)]
            return ans, alp
        if mode == 2:
            return [[opt[k]] for opt, k in zip(cp.options, first)], [[float(v)] for v in best]
        ans, alp = [], []
        sel_host = lp_host = None
        for q, opts in enumerate(cp.options):
            if count[q] == 1:
                ans.append([opts[first[q]]])
                alp.append([float(best[q])])
            elif count[q] == 0:
                ans.append([])
                alp.append([])
            else:   # exact ties: every option attaining the maximum (find_max_ind)
                if sel_host is None:
                    sel_host, lp_host = sel.cpu().numpy(), lp.cpu().numpy()
                a, b = int(cp.seg[q]), int(cp.seg[q + 1])
                keep = sel_host[a:b] != 0
                ans.append([o for o, k in zip(opts, keep) if k])
                alp.append([float(v) for v, k in zip(lp_host[a:b], keep) if k])
        return ans, alp

    def _answers(self, cp, lp):
        """Host-side answer lists from the log-probabilities (the same semantics evaluated with numpy; the tests hold
        the device version to it) (reference: batch_gqa_ops.py:222-225, 404-407, 744-748,
        util.find_max_ind util.py:64-66)."""
        if cp.kind == STATEMENT:
            return [[n] for n in cp.names], []
        if cp.kind == BINARY:
            p = np.exp(lp.astype(np.float32))
            ans = [['yes'] if v > 0.5 else ['no'] for v in p]
            alp = [[math.log(float(v))] if v > 0.5 else [math.log(1.0 - float(v))] for v in p]
            return ans, alp
        if cp.terminal == 'compare':
            ans, alp = [], []
            for q, opt in enumerate(cp.options):
                pair = lp[2 * q:2 * q + 2]
                k = int(np.argmax(pair))
                ans.append([opt[k]])
                alp.append([float(pair[k])])
            return ans, alp
        ans, alp = [], []
        p = np.exp(lp.astype(np.float32))
        for q, opts in enumerate(cp.options):
            a, b = int(cp.seg[q]), int(cp.seg[q + 1])
            seg = p[a:b]
            keep = (seg == seg.max()) & (seg > self._likelihood_threshold)
            ans.append([o for o, k in zip(opts, keep) if k])
            alp.append([float(v) for v, k in zip(lp[a:b], keep) if k])
        return ans, alp


def targets_of(cp, answers):
    """Loss targets of VQATrainer._compute_loss (reference: nsvqa/train/trainer.py:185-230) as a float array."""
    if cp.kind == BINARY:
        return np.asarray([1.0 if a in YES else 0.0 for a in answers], dtype=np.float32)
    if cp.kind == QUERY:
        return np.asarray([1.0 if a == o else 0.0 for a, op in zip(answers, cp.options) for o in op], dtype=np.float32)
    return np.zeros(cp.lp_num, dtype=np.float32)


class FusedTrainStep(object):
    """One training step of VQATrainer._train_batch (reference: nsvqa/train/trainer.py:429-442) without leaving
    libdfol_b200: scene + programs forward, loss, backward, (NCCL gradient all-reduce), clip_grad_norm_, Adam.

    All trainable oracle parameters are re-homed into one flat fp32 bucket (gradients likewise), so the
    data-parallel exchange is a single all-reduce and clip/Adam are one kernel each over the bucket.
    """

    def __init__(self, interpreter, lr=1e-4, weight_decay=1e-10, clip_norm=0.65, betas=(0.9, 0.999), eps=1e-8,
                 process_group=None, l1_lambda=0.0):
        self.interp = interpreter
        # config key ``l1_lambda`` (trainer.py:257-259): loss += l1_lambda * ||theta||_1 / numel over the trainable
        # parameters, scaled by 1 / batch with the rest of the loss (:434-435)
        self.l1_lambda = float(l1_lambda)
        self._last_total = 1
        self.engine = interpreter._engine
        self.lr, self.wd, self.clip = lr, weight_decay, clip_norm
        self.betas, self.eps = betas, eps
        self.group = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        oracle_params = interpreter.oracle_parameters()
        # the optimizer sees the trainable parameters only (VQATrainer.get_parameter_list); frozen oracle tensors
        # (sample_config.yaml freezes all four networks) stay out of the bucket and get scratch gradient buffers
        self.attention_params = [p for p in interpreter.attention_parameters() if p.requires_grad]
        # Bucket order = [gradients that are final LATE in the backward pass | those that are final EARLY]: the
        # first-layer / featurizer weights (and the featurizer bias, and the attention networks) are written by the last
        # kernels of the step; every other oracle gradient is complete once the table layers, the second layers and the
        # pair hidden layer have run.  With several ranks the EARLY segment is all-reduced on a communication stream
        # while the remaining backward kernels run (reference: the reduce-add of nn.DataParallel, data_parallel.py:54-57)
        w = interpreter._weights
        late_ids = {id(w.feat.weight), id(w.feat.bias), id(w.attr[0].weight), id(w.rel[0].weight)}
        trainable = [p for p in oracle_params if p.requires_grad]
        late = [p for p in trainable if id(p) in late_ids] + self.attention_params
        early = [p for p in trainable if id(p) not in late_ids]
        params = late + early
        assert params, 'nothing to train'
        self.params = params
        self.oracle_trainable = any(p.requires_grad for p in oracle_params)
        dev = params[0].device
        self.early_offset = sum(p.numel() for p in late)
        self.early_numel = sum(p.numel() for p in early)
        self._early_work = None
        self._comm_stream = None
        self.overlap = os.environ.get('DFOL_AR_OVERLAP', '1') != '0'
        self.bucket = FlatBucket(params)
        self.flat, self.flat_grad, self.grads = self.bucket.flat, self.bucket.flat_grad, dict(self.bucket.grads)
        for p in oracle_params:
            if not p.requires_grad and self.oracle_trainable:
                self.grads[id(p)] = torch.zeros_like(p, dtype=torch.float32)
        for p in interpreter.attention_parameters():
            if not p.requires_grad:  # frozen attention networks: scratch buffers for the backward reductions
                self.grads[id(p)] = torch.zeros_like(p, dtype=torch.float32)
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.step_count = 0
        self.scalars = torch.zeros(2, device=dev, dtype=torch.float32)  # [loss, grad sumsq]

    def _targets(self, pb, cp, dev):
        t = getattr(pb, '_dfol_targets', None)
        if t is None or t.device != dev:
            t = torch.from_numpy(targets_of(cp, pb._answers)).to(dev, non_blocking=True)
            pb._dfol_targets = t
            host = getattr(pb, '_dfol_host', None)
            if host is not None:
                host._dfol_targets = t
        return t

    def forward_backward(self, program_batch_list, global_question_num=None):
        """Accumulates gradients of loss / global_question_num into the flat bucket; returns the device loss scalar
        (local share)."""
        interp = self.interp
        dropout = interp._dropout_for(True)
        total = global_question_num or sum(pb.batch_size() for pb in program_batch_list)
        scale = 1.0 / float(total)
        self._last_total = total
        self.flat_grad.zero_()
        self.scalars.zero_()
        for k, pb in enumerate(program_batch_list):
            drop_k = None if dropout is None else (dropout[0], dropout[1] + k)
            feats = interp._features_of(pb)
            dev = feats[1].device if isinstance(feats, tuple) else feats.device
            st = capi.stream_ptr(dev)
            cp = interp.compiled(pb, False)
            layout = SceneLayout.of_compiled(cp, dev, dense_pairs=drop_k is not None)
            scene = self.engine.build_scene(feats, layout, keep_for_backward=self.oracle_trainable, cp=cp,
                                            dropout=drop_k)
            att = interp.modulator(cp)
            mod_ctx = None
            if att is not None:
                scene.mods, mod_ctx = att.forward(cp)
                scene.d_mods = torch.zeros_like(scene.mods)
            lp, tape = self.engine.run_programs(cp, scene, save_tape=True)
            target = self._targets(pb, cp, dev)
            d_lp = torch.empty_like(lp)
            seg = self.engine.upload_programs(cp, dev).get('seg')
            call('dfol_loss_fwd_bwd', ptr(lp), ptr(target), ptr(seg), cp.question_num, cp.lp_num, cp.kind, scale,
                 ptr(self.scalars), ptr(d_lp), st)
            if self.oracle_trainable:
                last = k == len(program_batch_list) - 1
                hook = self._reduce_early if (last and self.world > 1 and self.overlap and self.early_numel) else None
                self.engine.backward(cp, scene, tape, d_lp, self.grads, early_hook=hook)
            else:
                self.engine.program_backward(cp, scene, tape, d_lp)
            if mod_ctx is not None and self.attention_params:
                att.backward(mod_ctx, scene.d_mods, self.grads)
        return self.scalars[0]

    def _reduce_early(self):
        """Called by the backward pass at the point where the EARLY segment of the bucket is final: its all-reduce starts
        on the communication stream, behind an event of the compute stream, and overlaps the rest of the backward."""
        dev = self.flat.device
        main = torch.cuda.current_stream(dev)
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(dev)
        ev = torch.cuda.Event()
        ev.record(main)
        self._comm_stream.wait_event(ev)
        with torch.cuda.stream(self._comm_stream):
            self._early_work = torch.distributed.all_reduce(self.flat_grad[self.early_offset:],
                                                            op=torch.distributed.ReduceOp.SUM, group=self.group,
                                                            async_op=True)

    def reduce_gradients(self):
        """Sum of the flat gradient bucket over the data-parallel ranks (NCCL; replaces the reduce-add of
        nn.DataParallel, reference nn/interpreter/data_parallel.py:54-57): one all-reduce, or -- when the backward pass
        started the early segment on the communication stream -- the late segment now and a wait for the early one."""
        if self.world <= 1:
            return
        if self._early_work is not None:
            if self.early_offset:
                torch.distributed.all_reduce(self.flat_grad[:self.early_offset], op=torch.distributed.ReduceOp.SUM,
                                             group=self.group)
            self._early_work.wait()   # the compute stream waits for the communication stream
            self._early_work = None
        else:
            self.bucket.all_reduce(self.group)

    def optimizer_step(self):
        dev = self.flat.device
        st = capi.stream_ptr(dev)
        self.reduce_gradients()
        self.step_count += 1
        if self.l1_lambda > 0.0:
            # added ONCE, after the data-parallel sum; every rank reports 1 / world of the term in its loss share
            coef = self.l1_lambda / (float(self.flat.numel()) * float(self._last_total))
            call('dfol_l1_regularize', ptr(self.flat), ptr(self.flat_grad), self.flat.numel(), coef,
                 coef / float(self.world), ptr(self.scalars), st)
        call('dfol_sumsq', ptr(self.flat_grad), self.flat_grad.numel(), ptr(self.scalars[1:]), st)
        call('dfol_adam_step', ptr(self.flat), ptr(self.flat_grad), ptr(self.m), ptr(self.v), self.flat.numel(),
             ptr(self.scalars[1:]), self.clip, self.lr, self.betas[0], self.betas[1], self.eps, self.wd,
             self.step_count, st)

    def step(self, program_batch_list, global_question_num=None):
        loss = self.forward_backward(program_batch_list, global_question_num)
        self.optimizer_step()
        return loss
