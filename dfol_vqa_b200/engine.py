"""Host orchestration of the reasoning path: scene layout, scene build (visual oracle), program execution, and the
hand-derived backward pass, all as launches of libdfol_b200 kernels on the current CUDA stream.

Reference path being replaced: BatchInterpreterBase.build_scene + forward (nsvqa/nn/interpreter/
batch_base_interpreter.py:45-183) and the autograd backward of everything it calls.  PyTorch is used for device
memory (torch.empty / zeros), the stream, and index plumbing only.
"""

import numpy as np
import torch

from . import capi
from .capi import K, call, ptr

DEFAULT_LL = -30.0
# dropout sites (one independent mask each; rows x logical columns)
DROP_FEATURES, DROP_ATTR_IN, DROP_ATTR_HIDDEN, DROP_REL_IN, DROP_REL_HIDDEN, DROP_EMB_ATTR, DROP_EMB_REL = range(7)


def _roundup(x, m):
    return (x + m - 1) // m * m


def layout_tables(counts, concept_num, relation_num, pair_mask=None):
    """Host (numpy) row / offset tables of one program batch + its scalar sizes.

    ``pair_mask`` (bool per image, tensor-core mode with demand-driven relation slots): images whose program reads no
    relation likelihood get NO pair rows -- the pair-level chain (hidden layer, layer 2, slot columns and their
    backward) runs over the images that need it only.  No output depends on the skipped rows and their gradient
    contribution is exactly zero (the reference evaluates them and never reads them, classifier_oracle.py:154)."""
    n = np.asarray(counts, dtype=np.int64)
    assert n.ndim == 1 and n.size > 0 and int(n.min()) >= 1, 'every image needs at least one object'
    npair = n if pair_mask is None else np.where(np.asarray(pair_mask, dtype=bool), n, 0)
    a_stride = (n + 3) // 4 * 4
    r_stride = (n * n + 3) // 4 * 4
    attr_blk = concept_num * np.concatenate([[0], np.cumsum(a_stride)])
    rel_blk = relation_num * np.concatenate([[0], np.cumsum(r_stride)])
    arrays = {
        'img_n': n.astype(np.int32), 'img_np': npair.astype(np.int32), 'img_nn': (npair * npair).astype(np.int32),
        'obj_row': np.concatenate([[0], np.cumsum(n)]).astype(np.int32),
        'pair_row': np.concatenate([[0], np.cumsum(npair * npair)]).astype(np.int32),
        'attr_stride': a_stride.astype(np.int32), 'rel_stride': r_stride.astype(np.int32),
        'attr_blk': attr_blk[:-1].astype(np.int64), 'rel_blk': rel_blk[:-1].astype(np.int64),
        'obj_img': np.repeat(np.arange(n.size, dtype=np.int32), n),
        # 128-row tiles of the pair rows, per image (persistent tcgen05 kernels walk (image, tile) pairs)
        'pair_tile': np.concatenate([[0], np.cumsum((npair * npair + 127) // 128)]).astype(np.int32),
    }
    meta = {'counts': [int(v) for v in n], 'B': int(n.size), 'T': int(n.sum()), 'P': int((npair * npair).sum()),
            'max_n': int(n.max()), 'max_np': int(npair.max()), 'attr_size': int(attr_blk[-1]),
            'rel_size': int(rel_blk[-1]), 'masked': pair_mask is not None,
            'pair_tiles': int(((npair * npair + 127) // 128).sum()),
            'pair_images': int((npair > 0).sum())}
    assert meta['P'] < 2 ** 31 and meta['T'] < 2 ** 31
    return arrays, meta


class SceneLayout(object):
    """Row/offset tables of one program batch (see the layout comment in include/dfol_b200.h)."""

    _cache = {}
    _cache_bytes = 0
    _CACHE_LIMIT = 64 << 20

    def __init__(self, counts, concept_num, relation_num, device, pair_mask=None, tables=None):
        """``tables``: (device views, meta) of tables that were already uploaded with the batch's packed blob
        (compiler._pack_tables) -- no copy is issued then."""
        if tables is None:
            arrays, meta = layout_tables(counts, concept_num, relation_num, pair_mask)
            views = {k: torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=True)
                     for k, a in arrays.items()}
        else:
            views, meta = tables
        self.__dict__.update(meta)
        self.device = device
        for k, v in views.items():
            setattr(self, k, v)
        self._pair_img = None
        self.nbytes = sum(int(v.numel()) * v.element_size() for v in views.values())

    @property
    def pair_img(self):
        """image of every pair row (P int32; only the fp32 / dropout / fused-inference paths read it: built on demand)"""
        if self._pair_img is None:
            img = torch.arange(self.B, device=self.device, dtype=torch.int32)
            self._pair_img = torch.repeat_interleave(img, self.img_nn.long(), output_size=self.P)
            self.nbytes += self.P * 4
        return self._pair_img

    @classmethod
    def get(cls, counts, concept_num, relation_num, device, pair_mask=None):
        """Layout built from the object counts alone (tests, tools, callers without a compiled batch).  The cache is
        bounded by bytes; batches that went through the compiler carry their layout tables in their own packed blob
        (``of_compiled``) and never come here."""
        mask_key = None if pair_mask is None else tuple(bool(m) for m in pair_mask)
        key = (tuple(int(c) for c in counts), concept_num, relation_num, str(device), mask_key)
        hit = cls._cache.get(key)
        if hit is None:
            if cls._cache_bytes > cls._CACHE_LIMIT or len(cls._cache) > 64:
                cls._cache.clear()
                cls._cache_bytes = 0
            hit = cls(counts, concept_num, relation_num, device, pair_mask)
            cls._cache[key] = hit
            cls._cache_bytes += hit.nbytes + 4 * hit.P
        return hit

    @classmethod
    def of_compiled(cls, cp, device, dense_pairs=False):
        """Layout of a compiled batch from the tables packed into its blob at compile (= collate) time: lives and dies
        with the batch's device tables, costs no extra copy and no host synchronisation.  ``dense_pairs``: pair rows for
        every image (dropout path), whatever the programs read."""
        from .compiler import upload_tables
        cache = upload_tables(cp, device)
        key = 'layout_dense' if dense_pairs else 'layout'
        hit = cache.get(key)
        if hit is None:
            meta = dict(cp.layout_meta)
            views = dict(cache['lay'])
            if dense_pairs and meta['masked']:
                views['img_np'], views['img_nn'], views['pair_row'] = views['img_n'], views['img_nn_d'], views['pair_row_d']
                views['pair_tile'] = views['pair_tile_d']
                meta.update(P=meta['P_dense'], max_np=meta['max_n'], masked=False, pair_images=meta['B'],
                            pair_tiles=meta['pair_tiles_dense'])
            for k in ('img_nn_d', 'pair_row_d', 'pair_tile_d'):
                views.pop(k, None)
            hit = cache[key] = cls(None, None, None, device, tables=(views, meta))
        return hit


# fp32 parity mode: contractions large enough for the tensor cores run there as split-bf16 GEMMs (x = h + m + l, six
# product terms, fp32 TMEM accumulation: csrc/gemm_tcgen05.cu dfol_split3_bf16); small ones, strided operands the TMA
# maps cannot describe, and DFOL_FP32_SIMT=1 use the SIMT kernel (exact fp32 FMA chain).
_FP32_TC_MIN_WORK = 1 << 24


def _fp32_tc_enabled():
    import os
    return os.environ.get('DFOL_FP32_SIMT', '0') != '1'


def _split3(x, pattern, stacked, Kp, st):
    """Concatenated three-way bf16 split of the row-major fp32 matrix view ``x`` (rows, cols)."""
    rows, cols = x.shape
    if stacked:
        ld = _roundup(cols, 8)
        out = torch.empty(6 * rows, ld, device=x.device, dtype=torch.bfloat16)   # (pad columns are written as zeros)
    else:
        ld = 6 * Kp
        out = torch.empty(rows, ld, device=x.device, dtype=torch.bfloat16)
    call('dfol_split3_bf16', ptr(x), x.stride(0), rows, cols, ptr(out), ld, Kp, pattern, int(stacked), st)
    return out


def _gemm_f32_tc(A, B, C, bias, act, accumulate, split_k, table, st):
    """Tensor-core evaluation of gemm_f32's contract; returns False when the operand layout is not one of the three forms
    the path uses (NT: activations x weight^T, NN: gradient x weight, TN: gradient^T x activations)."""
    M, Kd = A.shape
    N = B.shape[1]
    a_row = A.stride(1) == 1 and A.stride(0) >= Kd
    b_col = B.stride(0) == 1 and B.stride(1) >= Kd          # B = W^T view of a row-major W [N, K]
    b_row = B.stride(1) == 1 and B.stride(0) >= N           # B row-major [K, N]
    a_col = A.stride(0) == 1 and A.stride(1) >= M           # A = X^T view of a row-major X [K, M]
    if a_col and b_row and table is None and bias is None and act == K.ACT_NONE:
        # TN (weight gradient): C[M, N] (+)= X^T . B, reduction over the Kd rows of X and B -> MN-major wgrad kernel
        X = A.t()
        if capi.trace is not None:
            capi.next_meta = {'tag': 'gemm_f32_tc_wgrad[%dx%dx%d]' % (M, N, Kd), 'flops': 12.0 * M * N * Kd}
        if not accumulate and split_k == 1:
            C.zero_()
        xa, xb = _split3(X, 0, True, 0, st), _split3(B, 1, True, 0, st)
        assert C.stride(1) == 1
        call('dfol_gemm_bf16_tc_wgrad', ptr(xa), xa.stride(0), ptr(xb), xb.stride(0), ptr(C), C.stride(0), M, N,
             6 * Kd, st)
        return True
    if not a_row:
        return False
    if b_col:
        W = B.t()                                            # row-major [N, K]
    elif b_row:
        W = B.t().contiguous()                               # NN (input gradient): small weight matrix, transposed once
    else:
        return False
    Kp = _roundup(Kd, 64)
    xa, xb = _split3(A, 0, False, Kp, st), _split3(W, 1, False, Kp, st)
    if table is None:
        out = C if not accumulate else torch.empty(M, N, device=C.device, dtype=torch.float32)
        assert out.stride(1) == 1
        ldc, store, maps, diag = out.stride(0), 0, (None, None, None, None, None), 0.0
    else:
        out, ldc, store = C, 0, 1
        maps = (ptr(table['row_img']), ptr(table['img_row']), ptr(table['img_blk']), ptr(table['img_stride']),
                ptr(table.get('img_n')))
        diag = table.get('diag', DEFAULT_LL)
    if capi.trace is not None:
        capi.next_meta = {'tag': 'gemm_f32_tc[%dx%dx%d]%s' % (M, N, Kd, ' table' if store else ''),
                          'flops': 12.0 * M * N * Kd}
    call('dfol_gemm_bf16_tc_exact', ptr(xa), 6 * Kp, ptr(xb), 6 * Kp, ptr(out), ldc, N, ptr(bias), M, N, 6 * Kp, act,
         store, maps[0], maps[1], maps[2], maps[3], maps[4], diag, st)
    if accumulate and table is None:
        C += out
    return True


def gemm_f32(A, B, C, bias=None, act=K.ACT_NONE, accumulate=False, split_k=1, mul_src=None, mul_mode=K.MUL_NONE,
             table=None, stream=None):
    """C = epilogue(A @ B) with A (M,K) and B (K,N) arbitrary-stride fp32 views; see dfol_gemm_f32."""
    M, Kd = A.shape
    Kb, N = B.shape
    assert Kd == Kb
    if (mul_src is None and float(M) * N * Kd >= _FP32_TC_MIN_WORK and min(M, N) >= 16 and _fp32_tc_enabled()
            and M < 65535 * 128):
        st = stream if stream is not None else capi.stream_ptr(A.device)
        if _gemm_f32_tc(A, B, C, bias, act, accumulate, split_k, table, st):
            return
    if table is None:
        assert C.shape == (M, N) and C.stride(1) == 1
        ldc, store, maps, diag = C.stride(0), 0, (None, None, None, None, None), 0.0
    else:
        ldc, store = 0, 1
        maps = (ptr(table['row_img']), ptr(table['img_row']), ptr(table['img_blk']), ptr(table['img_stride']),
                ptr(table.get('img_n')))
        diag = table.get('diag', DEFAULT_LL)
    if capi.trace is not None:
        capi.next_meta = {'tag': 'gemm_f32[%dx%dx%d]%s' % (M, N, Kd, ' table' if store else ''),
                          'flops': 2.0 * M * N * Kd}
    call('dfol_gemm_f32', ptr(A), A.stride(0), A.stride(1), ptr(B), B.stride(0), B.stride(1), ptr(C), ldc, ptr(bias),
         M, N, Kd, act, int(accumulate), split_k, ptr(mul_src), 0 if mul_src is None else mul_src.stride(0), mul_mode,
         store, maps[0], maps[1], maps[2], maps[3], maps[4], diag, stream)


def _split_for(k):
    return int(max(1, min(512, k // 2048)))


class OracleWeights(object):
    """Views of the 12 parameter tensors (reference key order)."""

    def __init__(self, featurizer_layers, attribute_layers, relation_layers, embedding_layer):
        assert len(featurizer_layers) == 1, 'featurizer_layers_config must be [] (single Linear + Sigmoid)'
        assert len(attribute_layers) >= 1 and len(relation_layers) >= 1
        self.feat = featurizer_layers[0]
        self.attr = attribute_layers
        self.rel = relation_layers
        self.emb = embedding_layer
        assert self.emb.bias is not None, 'freeze_embedding_bias is not supported by the fused path'

    def parameters(self):
        out = [self.feat.weight, self.feat.bias]
        for l in self.attr + self.rel:
            out += [l.weight, l.bias]
        out += [self.emb.weight, self.emb.bias]
        return out


class Scene(object):
    """Device tensors of one built scene (tables + the activations saved for backward)."""
    pass


class ReasoningEngine(object):

    def __init__(self, weights, relation_index, gemm_mode='fp32'):
        self.w = weights
        self.rel_index_host = list(relation_index)
        self._rel_index = None
        self.gemm_mode = gemm_mode
        if gemm_mode not in ('fp32', 'bf16'):
            raise ValueError('gemm_mode must be fp32 or bf16')
        from .engine_tc import TensorCorePath
        self.tc = TensorCorePath(self)

    def rel_index(self, device):
        if self._rel_index is None or self._rel_index.device != device:
            self._rel_index = torch.tensor(self.rel_index_host, dtype=torch.int64, device=device)
        return self._rel_index

    # ------------------------------------------------------------------------------------------ forward

    def build_scene(self, features, layout, keep_for_backward=True, cp=None, dropout=None):
        """Featurizer + attribute / relation tables (K1-K6 of SURVEY.md §2.1).  ``cp``: the compiled programs of the
        batch; when they carry relation slots only those relation columns are evaluated (tensor-core mode).
        ``dropout``: None, or (p, seed) -- training-mode dropout in front of every Linear (forward only: the oracle
        networks must be frozen, as in sample_config.yaml); see csrc/dropout_kernels.cu."""
        if self.gemm_mode == 'bf16':
            return self.tc.build_scene(features, layout, keep_for_backward, cp, dropout)
        capi.lib()
        w = self.w
        dev = features.device
        st = capi.stream_ptr(dev)
        T, width = features.shape
        D = width - 6
        assert T == layout.T and features.dtype == torch.float32 and features.stride(1) == 1
        F = w.feat.weight.shape[0]
        ldo = F + 4
        sc = Scene()
        sc.layout = layout
        sc.features = features

        def drop(x, cols, site):
            """x[:, :cols] *= mask of dropout site ``site`` (in place); no-op without dropout."""
            if dropout is not None:
                call('dfol_dropout_scale', ptr(x), x.stride(0), x.shape[0], cols, int(x.dtype == torch.bfloat16),
                     int(dropout[1]), site, float(dropout[0]), st)
            return x

        # featurizer: obj = [sigmoid(X Wf^T + b) | box position]
        obj = torch.empty(T, ldo, device=dev, dtype=torch.float32)
        x_in = features[:, :D] if dropout is None else drop(features[:, :D].contiguous(), D, DROP_FEATURES)
        gemm_f32(x_in, w.feat.weight.t(), obj[:, :F], w.feat.bias, K.ACT_SIGMOID, stream=st)
        sc.dropout = dropout
        sc.feat_in = x_in          # what the featurizer GEMM read (masked under dropout): operand of its wgrad
        keep = dropout is not None and keep_for_backward
        call('dfol_box_position', ptr(features), features.stride(0), D, ptr(obj), ldo, F, T, st)
        sc.obj = obj

        # attribute chain -> attribute table (all C concept columns)
        # under dropout every Linear reads a MASKED copy of the previous output: the unmasked outputs (attr_h) give
        # act'(h) in the backward pass, the masked inputs (attr_in) are the wgrad operands
        h = obj if dropout is None else drop(obj.clone(), ldo, DROP_ATTR_IN)
        sc.attr_h, sc.attr_in = [], [h]
        for i, layer in enumerate(w.attr):
            assert dropout is None or len(w.attr) == 2, 'dropout: one hidden layer per network (reference configs)'
            out = torch.empty(T, layer.weight.shape[0], device=dev, dtype=torch.float32)
            gemm_f32(h, layer.weight.t(), out, layer.bias, K.ACT_ELU if i < len(w.attr) - 1 else K.ACT_SIGMOID,
                     stream=st)
            sc.attr_h.append(out)
            h = drop(out.clone() if keep else out, out.shape[1], DROP_ATTR_HIDDEN if i == 0 else DROP_EMB_ATTR)
            sc.attr_in.append(h)
        C = w.emb.weight.shape[0]
        attr_ll = torch.empty(layout.attr_size, device=dev, dtype=torch.float32)
        obj_table = {'row_img': layout.obj_img, 'img_row': layout.obj_row, 'img_blk': layout.attr_blk,
                     'img_stride': layout.attr_stride}
        gemm_f32(h, w.emb.weight.t(), attr_ll, w.emb.bias, K.ACT_LOGSIGMOID, table=obj_table, stream=st)
        sc.attr_ll = attr_ll

        # relation chain: first layer through the U/V decomposition, then dense layers over all pair rows
        first = w.rel[0]
        H = first.weight.shape[0]
        assert first.weight.shape[1] == 2 * ldo + 4, 'relation network input must be 2*(F+4)+4'
        uv = torch.empty(T, 2 * H, device=dev, dtype=torch.float32)
        gemm_f32(obj, first.weight[:, :ldo].t(), uv[:, :H], stream=st)
        gemm_f32(obj, first.weight[:, ldo:2 * ldo].t(), uv[:, H:], stream=st)
        sc.uv = uv
        wg = first.weight[:, 2 * ldo:]
        h1 = torch.empty(layout.P, H, device=dev, dtype=torch.float32)
        act1 = K.ACT_ELU if len(w.rel) > 1 else K.ACT_SIGMOID
        if dropout is None:
            call('dfol_pair_hidden_fwd', ptr(uv), uv.stride(0), ptr(obj[:, F:]), ldo, ptr(wg), first.weight.stride(0),
                 ptr(first.bias), ptr(h1), H, H, act1, 0, ptr(layout.pair_row), ptr(layout.obj_row),
                 ptr(layout.img_n), layout.B, layout.max_n, st)
        else:
            # an independent mask per pair element breaks the U[s] + V[o] factorisation: the first layer runs on the
            # materialised, masked pair matrix as in the reference (batch_gqa_boxfeatures_pipeline.py:260-281)
            assert len(w.rel) == 2, 'dropout: one hidden layer per network (reference configs)'
            width = 2 * ldo + 4
            pm = torch.empty(layout.P, width, device=dev, dtype=torch.float32)
            call('dfol_pair_features_dropout', ptr(obj), ldo, ldo, F, ptr(pm), width, width, 0, ptr(layout.pair_row),
                 ptr(layout.obj_row), ptr(layout.img_n), ptr(layout.pair_img), layout.P, int(dropout[1]),
                 DROP_REL_IN, float(dropout[0]), st)
            gemm_f32(pm, first.weight.t(), h1, first.bias, act1, stream=st)
            sc.rel_in = [pm if keep else None]
            del pm
        sc.rel_h = [h1]
        h = h1 if dropout is None else drop(h1.clone() if keep else h1, H, DROP_REL_HIDDEN)
        if dropout is not None:
            sc.rel_in.append(h)
        for i, layer in enumerate(w.rel[1:], start=1):
            out = torch.empty(layout.P, layer.weight.shape[0], device=dev, dtype=torch.float32)
            gemm_f32(h, layer.weight.t(), out, layer.bias, K.ACT_ELU if i < len(w.rel) - 1 else K.ACT_SIGMOID,
                     stream=st)
            sc.rel_h.append(out)
            h = drop(out.clone() if keep else out, out.shape[1], DROP_EMB_REL)
            if dropout is not None:
                sc.rel_in.append(h)
        ridx = self.rel_index(dev)
        sc.w_rel = w.emb.weight.detach().index_select(0, ridx).contiguous()
        sc.b_rel = w.emb.bias.detach().index_select(0, ridx).contiguous()
        rel_ll = torch.empty(layout.rel_size, device=dev, dtype=torch.float32)
        pair_table = {'row_img': layout.pair_img, 'img_row': layout.pair_row, 'img_blk': layout.rel_blk,
                      'img_stride': layout.rel_stride, 'img_n': layout.img_n, 'diag': DEFAULT_LL}
        gemm_f32(h, sc.w_rel.t(), rel_ll, sc.b_rel, K.ACT_LOGSIGMOID, table=pair_table, stream=st)
        sc.rel_ll = rel_ll
        sc.rel_blk = layout.rel_blk
        return sc

    def upload_programs(self, cp, device):
        """Device copies of the tables of a CompiledPrograms (compiler.upload_tables: one packed copy, cached)."""
        from .compiler import upload_tables
        return upload_tables(cp, device)

    def run_programs(self, cp, scene, save_tape=True):
        """Executes the compiled programs; returns lp (cp.lp_num,) and the tape (or None).  ``scene.mods`` (optional,
        (cp.mod_rows, 4) fp32): attention-transfer modulations; program_backward then fills ``scene.d_mods``."""
        lay = scene.layout
        dev = scene.attr_ll.device
        st = capi.stream_ptr(dev)
        assert lay.max_n <= 128, 'the interpreter kernels support at most 128 objects per image'
        d = self.upload_programs(cp, dev)
        lp = torch.empty(max(cp.lp_num, 1), device=dev, dtype=torch.float32)
        n_instr = cp.instr.shape[0]
        stride = _roundup(lay.max_n, 4)
        tape = torch.empty(max(n_instr, 1) * stride, device=dev, dtype=torch.float32) if save_tape else None
        if capi.trace is not None:
            capi.next_meta = {'tag': 'program_fwd', 'bytes': cp.alg_bytes}
        fast = self.gemm_mode == 'bf16'
        # tensor-core mode: the probability table (scene.rel_p, written by the slot kernels) feeds the relate hops
        rel = (ptr(scene.rel_ll), ptr(getattr(scene, 'rel_p', None))) if fast else (ptr(scene.rel_ll),)
        call('dfol_program_fwd_fast' if fast else 'dfol_program_fwd', ptr(d['instr']), ptr(d['q_instr']),
             ptr(d['opts']), cp.question_num, ptr(scene.attr_ll), ptr(lay.attr_blk), ptr(lay.attr_stride), *rel,
             ptr(scene.rel_blk), ptr(lay.rel_stride), ptr(lay.img_n), ptr(getattr(scene, 'mods', None)), ptr(lp),
             ptr(tape), stride, st)
        return lp[:cp.lp_num], tape

    # ------------------------------------------------------------------------------------------ backward

    def _slice_tables(self, slices, B, device, key, cp):
        """Per-image grouped slice tables for dfol_table_layer_bwd* (packed at compile time, uploaded with the bytecode)."""
        return self.upload_programs(cp, device)[key]

    def program_backward(self, cp, scene, tape, d_lp):
        """Backward interpreter: compact gradient slices w.r.t. the raw attribute / relation table entries."""
        lay = scene.layout
        dev = scene.attr_ll.device
        st = capi.stream_ptr(dev)
        d = self.upload_programs(cp, dev)
        stride = _roundup(lay.max_n, 4)
        g_attr = torch.zeros(cp.g_attr_size, device=dev, dtype=torch.float32)
        g_rel = torch.zeros(cp.g_rel_size, device=dev, dtype=torch.float32)
        if capi.trace is not None:
            capi.next_meta = {'tag': 'program_bwd', 'bytes': 2.0 * cp.alg_bytes}
        fast = self.gemm_mode == 'bf16'
        rel = (ptr(scene.rel_ll), ptr(getattr(scene, 'rel_p', None))) if fast else (ptr(scene.rel_ll),)
        call('dfol_program_bwd_fast' if fast else 'dfol_program_bwd', ptr(d['instr']), ptr(d['q_instr']),
             ptr(d['opts']), cp.question_num, ptr(scene.attr_ll), ptr(lay.attr_blk), ptr(lay.attr_stride), *rel,
             ptr(scene.rel_blk), ptr(lay.rel_stride), ptr(lay.img_n), ptr(getattr(scene, 'mods', None)), ptr(d_lp),
             ptr(tape), stride, ptr(g_attr), ptr(g_rel), ptr(getattr(scene, 'd_mods', None)), st)
        return g_attr, g_rel

    def backward(self, cp, scene, tape, d_lp, grads, early_hook=None):
        """d loss / d parameters given d loss / d lp.  ``grads``: dict param tensor id -> fp32 grad tensor of the
        parameter's shape (accumulated into; callers zero them).  ``early_hook``: called (tensor-core mode) once every
        gradient except the first-layer / featurizer weights and the featurizer bias is final on the current stream."""
        if self.gemm_mode == 'bf16':
            return self.tc.backward(cp, scene, tape, d_lp, grads, early_hook)
        w = self.w
        lay = scene.layout
        dev = scene.attr_ll.device
        st = capi.stream_ptr(dev)
        g_attr, g_rel = self.program_backward(cp, scene, tape, d_lp)

        def G(p):
            return grads[id(p)]

        T, P = lay.T, lay.P
        obj = scene.obj
        F = w.feat.weight.shape[0]
        ldo = F + 4
        E = w.emb.weight.shape[1]
        d_obj = torch.zeros(T, ldo, device=dev, dtype=torch.float32)

        dropout = getattr(scene, 'dropout', None)

        def drop(x, cols, site):
            """gradient w.r.t. a masked tensor -> gradient w.r.t. the unmasked one: the same mask again (in place)"""
            call('dfol_dropout_scale', ptr(x), x.stride(0), x.shape[0], cols, 0, int(dropout[1]), site,
                 float(dropout[0]), st)
            return x

        # ---- attribute table layer (sparse slices) -> dense chain backward
        sa = self._slice_tables(cp.attr_slices, lay.B, dev, 'attr_slices', cp)
        if dropout is None:
            h_last = scene.attr_h[-1]
            d_h, is_dz = self._table_backward(
                g_attr, sa, scene.attr_ll, lay.attr_blk, lay.attr_stride, lay.obj_row, lay.img_n, lay.max_n, T,
                w.emb.weight, G(w.emb.weight), G(w.emb.bias), h_last, K.ACT_SIGMOID, st)
            self._mlp_backward(w.attr, [obj] + scene.attr_h, d_h, d_obj, grads, st, first_layer_input_grad=True,
                               d_out_is_dz=is_dz)
        else:
            # the table layer read the MASKED last activation: un-fused (no act'), mask the gradient, then the chain
            d_h, _ = self._table_backward(
                g_attr, sa, scene.attr_ll, lay.attr_blk, lay.attr_stride, lay.obj_row, lay.img_n, lay.max_n, T,
                w.emb.weight, G(w.emb.weight), G(w.emb.bias), scene.attr_in[-1], K.ACT_NONE, st)
            drop(d_h, E, DROP_EMB_ATTR)
            self._mlp_backward(w.attr, [obj] + scene.attr_h, d_h, d_obj, grads, st, first_layer_input_grad=True,
                               ins=scene.attr_in, mask=lambda x, i: drop(x, x.shape[1],
                                                                         (DROP_ATTR_IN, DROP_ATTR_HIDDEN)[i]))

        # ---- relation table layer -> dense layers -> pair hidden layer
        sr = self._slice_tables(cp.rel_slices, lay.B, dev, 'rel_slices', cp)
        if sr['count']:
            nR = scene.w_rel.shape[0]
            h_last = scene.rel_h[-1]
            dw_rel = torch.zeros(nR, E, device=dev, dtype=torch.float32)
            db_rel = torch.zeros(nR, device=dev, dtype=torch.float32)
            fuse = K.ACT_SIGMOID if (len(w.rel) > 1 and dropout is None) else K.ACT_NONE
            if dropout is not None:
                h_last = scene.rel_in[-1]   # the table layer read the masked activation
            d_h, is_dz = self._table_backward(
                g_rel, sr, scene.rel_ll, lay.rel_blk, lay.rel_stride, lay.pair_row, lay.img_nn, lay.max_n ** 2, P,
                scene.w_rel, dw_rel, db_rel, h_last, fuse, st)
            ridx = self.rel_index(dev)
            G(w.emb.weight).index_add_(0, ridx, dw_rel)
            G(w.emb.bias).index_add_(0, ridx, db_rel)
            first = w.rel[0]
            H = first.weight.shape[0]
            gw1 = G(first.weight)
            if dropout is not None:
                # masked layers: layer 2 on the masked h1, layer 1 on the materialised masked pair matrix
                drop(d_h, E, DROP_EMB_REL)
                d_h1 = self._mlp_backward(w.rel[1:], scene.rel_h, d_h, None, grads, st, first_layer_input_grad=False,
                                          ins=scene.rel_in[1:], mask=lambda x, i: drop(x, x.shape[1], DROP_REL_HIDDEN))
                h1, pm = scene.rel_h[0], scene.rel_in[0]
                call('dfol_act_grad_mul', ptr(d_h1), d_h1.stride(0), ptr(h1), h1.stride(0), P, H, K.ACT_ELU, st)
                call('dfol_colsum', ptr(d_h1), d_h1.stride(0), P, H, ptr(G(first.bias)), st)
                sk = _split_for(P)
                gemm_f32(d_h1.t(), pm, gw1, accumulate=(sk == 1), split_k=sk, stream=st)
                d_pm = torch.empty(P, pm.shape[1], device=dev, dtype=torch.float32)
                gemm_f32(d_h1, first.weight, d_pm, stream=st)
                drop(d_pm, pm.shape[1], DROP_REL_IN)
                call('dfol_pair_features_bwd', ptr(d_pm), d_pm.stride(0), 0, ldo, ptr(d_obj), ldo, None, 0,
                     ptr(lay.pair_row), ptr(lay.obj_row), ptr(lay.img_n), ptr(lay.obj_img), T, st)
                sr = {'count': 0}  # the factored first-layer backward below is skipped
        if sr['count']:
            # dense layers above the pair hidden layer
            d_h1 = self._mlp_backward(w.rel[1:], scene.rel_h, d_h, None, grads, st, first_layer_input_grad=False,
                                      d_out_is_dz=is_dz and len(w.rel) > 1)
            duv = torch.zeros(T, 2 * H, device=dev, dtype=torch.float32)
            act1 = K.ACT_ELU if len(w.rel) > 1 else K.ACT_SIGMOID
            call('dfol_pair_hidden_bwd', ptr(d_h1), d_h1.stride(0), ptr(scene.rel_h[0]), scene.rel_h[0].stride(0),
                 ptr(obj[:, F:]), ldo, ptr(duv), duv.stride(0), ptr(gw1[:, 2 * ldo:]), gw1.stride(0),
                 ptr(G(first.bias)), H, act1, ptr(lay.pair_row), ptr(lay.obj_row), ptr(lay.img_n), lay.B, st)
            sk = _split_for(T)
            gemm_f32(duv[:, :H].t(), obj, gw1[:, :ldo], accumulate=(sk == 1), split_k=sk, stream=st)
            gemm_f32(duv[:, H:].t(), obj, gw1[:, ldo:2 * ldo], accumulate=(sk == 1), split_k=sk, stream=st)
            gemm_f32(duv[:, :H], first.weight[:, :ldo], d_obj, accumulate=True, stream=st)
            gemm_f32(duv[:, H:], first.weight[:, ldo:2 * ldo], d_obj, accumulate=True, stream=st)

        # ---- featurizer: d pre = d obj[:, :F] * f (1 - f)
        call('dfol_act_grad_mul', ptr(d_obj), ldo, ptr(obj), ldo, T, F, K.ACT_SIGMOID, st)
        call('dfol_colsum', ptr(d_obj), ldo, T, F, ptr(G(w.feat.bias)), st)
        D = scene.features.shape[1] - 6
        sk = _split_for(T)
        x_in = scene.features[:, :D] if dropout is None else scene.feat_in
        gemm_f32(d_obj[:, :F].t(), x_in, G(w.feat.weight), accumulate=(sk == 1), split_k=sk, stream=st)

    def _table_backward(self, g, tabs, ll, blk, stride, row0, img_rows, max_rows, rows_total, W, dW, db, h_last,
                        fuse_act, st):
        """Backward of a table layer LL = logsigmoid(h_last W^T + b) from the compact program gradient slices.

        Returns (d, is_dz): d = d loss / d h_last, already multiplied by ``fuse_act``'(h_last) when is_dz is true.
        Few slices per image (relations, binary programs): one fused pass.  Many (attribute option lists): scatter to
        a dense d logits matrix and use GEMMs."""
        dev = ll.device
        E = W.shape[1]
        if tabs['count'] == 0:
            return torch.zeros(rows_total, E, device=dev, dtype=torch.float32), False
        if tabs['max_per_image'] <= 8:
            d = torch.empty(rows_total, E, device=dev, dtype=torch.float32)
            call('dfol_table_layer_bwd_fused', ptr(g), ptr(tabs['goff']), ptr(tabs['col']), ptr(tabs['col']),
                 ptr(tabs['img_slice']), len(tabs['img_slice']) - 1, max_rows, ptr(ll), ptr(blk), ptr(stride),
                 ptr(row0), ptr(img_rows), ptr(W), W.stride(0), ptr(h_last), h_last.stride(0), E, fuse_act, ptr(d),
                 d.stride(0), E, 0, ptr(dW), ptr(db), st)
            return d, fuse_act != K.ACT_NONE
        C = W.shape[0]
        dz = torch.zeros(rows_total, C, device=dev, dtype=torch.float32)
        call('dfol_table_grad_dense', ptr(g), ptr(tabs['goff']), ptr(tabs['col']), ptr(tabs['img']), tabs['count'],
             ptr(ll), ptr(blk), ptr(stride), ptr(row0), ptr(img_rows), ptr(dz), C, st)
        call('dfol_colsum', ptr(dz), C, rows_total, C, ptr(db), st)
        sk = _split_for(rows_total)
        gemm_f32(dz.t(), h_last, dW, accumulate=(sk == 1), split_k=sk, stream=st)
        d = torch.empty(rows_total, E, device=dev, dtype=torch.float32)
        gemm_f32(dz, W, d, stream=st)
        return d, False

    def _mlp_backward(self, layers, acts, d_out, d_in_accum, grads, st, first_layer_input_grad, d_out_is_dz=False,
                      ins=None, mask=None):
        """Backward through [Linear+act]* given d loss / d (last activation OUTPUT).

        ``acts`` = [input, h_1, ..., h_L] (saved outputs), ``layers`` the L Linear modules. Returns d loss / d input
        (accumulated into ``d_in_accum`` if given, else a fresh tensor) -- or d_out itself when L == 0.
        Dropout: ``ins[i]`` is what Linear i really read (the masked copy of acts[i]) and ``mask(x, i)`` multiplies the
        gradient w.r.t. that masked input by the same mask in place.
        """
        L = len(layers)
        if L == 0:
            return d_out
        assert len(acts) == L + 1
        d_h = d_out
        for i in range(L - 1, -1, -1):
            layer = layers[i]
            h_out, h_in = acts[i + 1], (acts[i] if ins is None else ins[i])
            rows, width = h_out.shape
            act = K.ACT_SIGMOID if i == L - 1 else K.ACT_ELU
            # dZ = dH * act'(h) in place (already applied by the fused table-layer kernel for the last layer)
            if not (d_out_is_dz and i == L - 1):
                call('dfol_act_grad_mul', ptr(d_h), d_h.stride(0), ptr(h_out), h_out.stride(0), rows, width, act, st)
            call('dfol_colsum', ptr(d_h), d_h.stride(0), rows, width, ptr(grads[id(layer.bias)]), st)
            sk = _split_for(rows)
            gemm_f32(d_h.t(), h_in, grads[id(layer.weight)], accumulate=(sk == 1), split_k=sk, stream=st)
            if i > 0 or first_layer_input_grad or d_in_accum is None:
                if i == 0 and d_in_accum is not None and mask is None:
                    gemm_f32(d_h, layer.weight, d_in_accum, accumulate=True, stream=st)
                    d_h = d_in_accum
                else:
                    nxt = torch.empty(rows, h_in.shape[1], device=d_h.device, dtype=torch.float32)
                    gemm_f32(d_h, layer.weight, nxt, stream=st)
                    if mask is not None:
                        mask(nxt, i)
                    if i == 0 and d_in_accum is not None:
                        d_in_accum += nxt
                        nxt = d_in_accum
                    d_h = nxt
        return d_h
