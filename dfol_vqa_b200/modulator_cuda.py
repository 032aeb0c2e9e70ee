"""Attention-transfer calibrator, token side, on libdfol_b200 (no torch compute): the same two LSTM passes and output
layer as ``modulator.AttentionTransfer`` (the device-agnostic torch statement of the reference's modulator loops, which
the CPU tests hold to the recorded reference runs), executed with hand-written kernels and a hand-derived backward.

Why not autograd over torch ops: a batch has ~10 op slots and ~3 500 predicate rows of 50-wide states -- the torch
version is ~1 000 tiny launches per step (forward + autograd), 6.8 ms of pure launch overhead against 2.1 ms for the
whole CUDA path of the calibrator-training step.  Here the input projections of ALL cells of the batch are ONE GEMM per
network (``features . W_ih^T``; the feature matrix depends only on the program and is cached with the bytecode), every
cell is one launch of the recurrent part (``dfol_lstm_cell_fwd``: W_hh in shared memory, expand / add / mask-gate fused),
every modulated sub-operator one launch of the output layer, and the parameter gradients are six GEMM-shaped reductions
over all cell rows at the end of the backward pass.

Reference: batch_base_interpreter.py:87-140, batch_base_ops.py:275-286, :407-467, :598-684 (see modulator.py for the
per-operator citations; the control flow below is the same, the tensor ops are kernel launches recorded on a tape).
"""

import numpy as np
import torch

from . import capi
from .capi import call, ptr
from .engine import gemm_f32
from .modulator import OPS_NUM, AttentionTransfer


class _State(object):
    """(h, c) of a pass, rows x S fp32; ``None`` tensors stand for the all-zero state."""

    __slots__ = ('h', 'c', 'dh', 'dc')

    def __init__(self, h=None, c=None):
        self.h, self.c, self.dh, self.dc = h, c, None, None


class NativeAttentionTransfer(object):

    def __init__(self, forward_network, backward_network, output_network, ontology):
        self.fwd, self.bwd, self.out = forward_network, backward_network, output_network
        self.lin = output_network[0]
        self.S = forward_network.hidden_size
        self.n_out = self.lin.weight.shape[0]
        assert self.S <= 64, 'attention_transfer_state_dim must be <= 64'
        self._torch = AttentionTransfer(forward_network, backward_network, output_network, ontology)
        self._tables = {}

    def parameters(self):
        return self._torch.parameters()

    def _embedding_table(self, dev):
        """(vocabulary, E) word-embedding table on the device (ontology.get_embeddings of every vocabulary entry: what
        ClassifierOracle.get_embedding returns per token, base_oracle.py:45-55)."""
        key = str(dev)
        hit = self._tables.get(key)
        if hit is None:
            ont = self._torch.ont
            emb = np.asarray(ont.get_embeddings(ont._vocabulary['idx_to_arg']), dtype=np.float32)
            hit = self._tables[key] = torch.from_numpy(emb).to(dev)
        return hit

    # ---- program-only data of a compiled batch (cached with the bytecode)

    def _plan(self, cp, dev):
        key = ('native', str(dev))
        hit = cp.mod_cache.get(key)
        if hit is not None:
            return hit
        # feature rows [one-hot operator | attribute/relation flag | word embedding] assembled on the device from the
        # compiler's per-row indices and the vocabulary embedding table (index plumbing; blank rows stay all-zero)
        from .compiler import upload_tables
        I = self.fwd.input_size
        table = self._embedding_table(dev)
        tabs = upload_tables(cp, dev)   # (already on the device when the batch was staged with its features)
        tok = tabs['mod_tok']
        R = tok.shape[0]
        live = (tok >= 0).to(torch.float32)           # (no nonzero(): nothing here synchronises with the host)
        feats = torch.zeros(R, I, device=dev, dtype=torch.float32)
        feats[torch.arange(R, device=dev), tabs['mod_opcol']] = live
        feats[:, OPS_NUM] = tabs['mod_relflag'] * live
        feats[:, OPS_NUM + 1:] = table[tok.clamp(min=0)] * live[:, None]
        base_of = {(slot_i, skey): (base, n) for slot_i, skey, n, base in cp.mod_plan}
        owners, masks = {}, {}
        for name, view in tabs.items():
            if name.startswith('mod_owner:'):
                _, i, k = name.split(':')
                owners[(int(i), k)] = view
            elif name.startswith('mod_mask:'):
                masks[int(name.split(':')[1])] = view
        hit = {'feats': feats, 'base': base_of, 'owners': owners, 'masks': masks}
        cp.mod_cache[key] = hit
        return hit

    # ---- forward

    def forward(self, cp):
        """Returns (mods (cp.mod_rows, 4) fp32, ctx for backward)."""
        dev = self.fwd.weight_ih.device
        st = capi.stream_ptr(dev)
        S, n_out, B = self.S, self.n_out, cp.question_num
        plan = self._plan(cp, dev)
        F_all = plan['feats']
        R = F_all.shape[0]
        x_f = torch.empty(R, 4 * S, device=dev, dtype=torch.float32)
        x_b = torch.empty(R, 4 * S, device=dev, dtype=torch.float32)
        gemm_f32(F_all, self.fwd.weight_ih.t(), x_f, self.fwd.bias_ih, stream=st)
        gemm_f32(F_all, self.bwd.weight_ih.t(), x_b, self.bwd.bias_ih, stream=st)
        saved = {'f': torch.zeros(R, 7 * S, device=dev, dtype=torch.float32),
                 'b': torch.zeros(R, 7 * S, device=dev, dtype=torch.float32)}
        mods = torch.empty(R, n_out, device=dev, dtype=torch.float32)
        cat = torch.zeros(R, 2 * S, device=dev, dtype=torch.float32)
        nets = {'f': (self.fwd, x_f), 'b': (self.bwd, x_b)}
        tape = []

        def zeros(rows):
            return torch.zeros(rows, S, device=dev, dtype=torch.float32)

        def cell(net, slot_key, state, add=None, owner=None, mask=None):
            base, rows = plan['base'][slot_key]
            mod, xp = nets[net]
            out = _State(torch.empty(rows, S, device=dev, dtype=torch.float32),
                         torch.empty(rows, S, device=dev, dtype=torch.float32))
            fb = None
            if mask is not None:
                fb = state
                if fb.h is None:  # zero fallback state: materialise it
                    fb.h, fb.c = zeros(rows), zeros(rows)
            call('dfol_lstm_cell_fwd', ptr(xp[base:]), xp.stride(0), ptr(mod.bias_hh), ptr(mod.weight_hh), S,
                 ptr(state.h), ptr(state.c), ptr(None if add is None else add.h), ptr(None if add is None else add.c),
                 ptr(owner), ptr(mask), ptr(None if fb is None else fb.h), ptr(None if fb is None else fb.c),
                 ptr(out.h), ptr(out.c), ptr(saved[net][base:]), rows, st)
            tape.append(('cell', net, base, rows, state, add if (add is not None and add.h is not None) else None,
                         owner, mask, fb, out))
            return out

        def out_layer(slot_key, fstate, bstate):
            base, rows = plan['base'][slot_key]
            call('dfol_mod_out_fwd', ptr(fstate.h), ptr(bstate.h), None, ptr(self.lin.weight), ptr(self.lin.bias), S,
                 n_out, ptr(mods[base:]), ptr(cat[base:]), rows, st)
            tape.append(('out', base, rows, fstate, bstate))

        def squeeze(state, owner):
            out = _State(zeros(B).index_add_(0, owner, state.h), zeros(B).index_add_(0, owner, state.c))
            tape.append(('squeeze', state, owner, out))
            return out

        descs = cp.mod_descs
        n = len(descs)
        fwd_state = {}

        def run(i, d, is_forward, inputs, gated):
            """One slot of one pass; returns the outgoing state (or a pair for two-input terminals)."""
            op = d['op']
            mask = plan['masks'].get(i) if gated else None
            if op == 'select':
                if d['select'] is None:
                    return _State() if is_forward else inputs[0]
                if is_forward:
                    fwd_state[(i, 'select')] = cell('f', (i, 'select'), _State())
                    return fwd_state[(i, 'select')]
                out_layer((i, 'select'), fwd_state[(i, 'select')], inputs[0])
                return inputs[0]  # the select's own backward cell feeds nothing (first slot of its branch)
            if op in ('relate', 'verify_rel', 'choose_rel'):
                state = inputs[0]
                owner = plan['owners'].get((i, 'relate'))
                if is_forward:
                    x = _State()
                    if d['select'] is not None:
                        x = fwd_state[(i, 'select')] = cell('f', (i, 'select'), _State())
                    new = cell('f', (i, 'relate'), state, add=x, owner=owner, mask=mask)
                    fwd_state[(i, 'relate')] = new
                    return new
                out_layer((i, 'relate'), fwd_state[(i, 'relate')], state)
                raw = cell('b', (i, 'relate'), state, owner=None, mask=None)
                if owner is not None:
                    raw = squeeze(raw, owner)
                if d['select'] is not None:
                    out_layer((i, 'select'), fwd_state[(i, 'select')], raw)
                if mask is None:
                    return raw
                return gate(raw, state, mask)
            if op in ('exist', 'end'):
                return inputs[0]
            if op in ('and', 'or'):
                return (inputs[0], inputs[1])
            fil = d.get('filter')
            if d.get('two'):
                if fil is None:
                    return (inputs[0], inputs[1])
                owner = plan['owners'].get((i, 'filter'))
                return (filter_like(i, 'filter0', is_forward, inputs[0], owner, None),
                        filter_like(i, 'filter1', is_forward, inputs[1], owner, None))
            if fil is None:
                return inputs[0]
            return filter_like(i, 'filter', is_forward, inputs[0], plan['owners'].get((i, 'filter')), mask)

        def gate(new, old, mask):
            """mask ? new : old (rows), recorded for the backward pass."""
            if old.h is None:
                old.h, old.c = zeros(new.h.shape[0]), zeros(new.h.shape[0])
            keep = (mask > 0).unsqueeze(1)
            out = _State(torch.where(keep, new.h, old.h), torch.where(keep, new.c, old.c))
            tape.append(('gate', new, old, keep, out))
            return out

        def filter_like(i, key, is_forward, state, owner, mask):
            if is_forward:
                new = cell('f', (i, key), state, owner=owner, mask=mask if owner is None else None)
                fwd_state[(i, key)] = new
                return new
            out_layer((i, key), fwd_state[(i, key)], state)
            new = cell('b', (i, key), state, mask=mask if owner is None else None)
            return squeeze(new, owner) if owner is not None else new

        trace = []
        for i, d in enumerate(descs):
            inputs = [trace[j] for j in d['deps']]
            trace.append(run(i, d, True, inputs, gated=(i < n - 1 and bool(inputs) and d['mask'] is not None)))
        consumers = [[] for _ in range(n)]
        for i, d in enumerate(descs):
            for j in d['deps']:
                consumers[j].append(i)
        two_in = isinstance(trace[-1], tuple)
        back = [None] * n
        for i in range(n - 1, -1, -1):
            d = descs[i]
            if len(consumers[i]) == 1:
                t = back[consumers[i][0]]
                inputs = [(t[1] if i == n - 2 else t[0]) if isinstance(t, tuple) else t]
            else:
                inputs = [_State(), _State()] if two_in else [_State()]
            x = run(i, d, False, inputs, gated=(bool(d['deps']) and d['mask'] is not None and i != n - 1))
            back[i] = x
        ctx = {'tape': tape, 'saved': saved, 'cat': cat, 'mods': mods, 'plan': plan, 'R': R}
        return mods[:cp.mod_rows], ctx

    # ---- backward

    def backward(self, ctx, d_mods, grads):
        """Accumulates d loss / d (attention-network parameters) into ``grads[id(param)]`` (fp32, parameter-shaped)."""
        dev = d_mods.device
        st = capi.stream_ptr(dev)
        S, n_out, R = self.S, self.n_out, ctx['R']
        saved, cat, mods, F_all = ctx['saved'], ctx['cat'], ctx['mods'], ctx['plan']['feats']
        d_mods = d_mods.contiguous().float()
        if d_mods.shape[0] < R:
            d_mods = torch.cat([d_mods, torch.zeros(R - d_mods.shape[0], n_out, device=dev)])
        dpre = {'f': torch.zeros(R, 4 * S, device=dev, dtype=torch.float32),
                'b': torch.zeros(R, 4 * S, device=dev, dtype=torch.float32)}
        dzo = torch.zeros(R, n_out, device=dev, dtype=torch.float32)
        w_hh = {'f': self.fwd.weight_hh, 'b': self.bwd.weight_hh}

        # gradient buffers of every state on the tape: ONE zero-filled pool, sliced (instead of ~100 tiny memsets)
        states = {}
        for rec in ctx['tape']:
            for obj in rec[1:]:
                if isinstance(obj, _State) and obj.h is not None:
                    states[id(obj)] = obj
        total = sum(o.h.shape[0] for o in states.values())
        pool = torch.zeros(2, max(total, 1), S, device=dev, dtype=torch.float32)
        off = 0
        for o in states.values():
            n = o.h.shape[0]
            o.dh, o.dc = pool[0, off:off + n], pool[1, off:off + n]
            off += n
        touched = set()

        def need(state, rows):
            touched.add(id(state))
            return state

        for rec in reversed(ctx['tape']):
            kind = rec[0]
            if kind == 'out':
                _, base, rows, fstate, bstate = rec
                need(fstate, rows)
                has_b = bstate.h is not None
                if has_b:
                    need(bstate, bstate.h.shape[0])
                call('dfol_mod_out_bwd', ptr(d_mods[base:]), ptr(mods[base:]), None, ptr(self.lin.weight), S, n_out,
                     ptr(dzo[base:]), ptr(fstate.dh), ptr(bstate.dh if has_b else None), rows, st)
            elif kind == 'cell':
                _, net, base, rows, state, add, owner, mask, fb, out = rec
                if id(out) not in touched:
                    continue  # nothing downstream depends on this cell
                has_in = state.h is not None
                if has_in:
                    need(state, state.h.shape[0])
                if add is not None:
                    need(add, add.h.shape[0])
                if fb is not None:
                    need(fb, fb.h.shape[0])
                call('dfol_lstm_cell_bwd', ptr(out.dh), ptr(out.dc), ptr(w_hh[net]), S, ptr(saved[net][base:]),
                     ptr(owner), ptr(mask), ptr(dpre[net][base:]), 4 * S, ptr(state.dh if has_in else None),
                     ptr(state.dc if has_in else None), ptr(None if add is None else add.dh),
                     ptr(None if add is None else add.dc), ptr(None if fb is None else fb.dh),
                     ptr(None if fb is None else fb.dc), rows, st)
            elif kind == 'squeeze':
                _, state, owner, out = rec
                if id(out) not in touched:
                    continue
                need(state, state.h.shape[0])
                state.dh += out.dh[owner]
                state.dc += out.dc[owner]
            elif kind == 'gate':
                _, new, old, keep, out = rec
                if id(out) not in touched:
                    continue
                need(new, new.h.shape[0])
                need(old, old.h.shape[0])
                zero = torch.zeros_like(out.dh)
                new.dh += torch.where(keep, out.dh, zero)
                new.dc += torch.where(keep, out.dc, zero)
                old.dh += torch.where(keep, zero, out.dh)
                old.dc += torch.where(keep, zero, out.dc)

        def G(p):
            return grads[id(p)]

        for net, mod in (('f', self.fwd), ('b', self.bwd)):
            dp = dpre[net]
            sk = max(1, min(16, R // 512))
            gemm_f32(dp.t(), F_all, G(mod.weight_ih), accumulate=(sk == 1), split_k=sk, stream=st)
            gemm_f32(dp.t(), saved[net][:, 6 * S:7 * S], G(mod.weight_hh), accumulate=(sk == 1), split_k=sk, stream=st)
            call('dfol_colsum', ptr(dp), dp.stride(0), R, 4 * S, ptr(G(mod.bias_ih)), st)
            call('dfol_colsum', ptr(dp), dp.stride(0), R, 4 * S, ptr(G(mod.bias_hh)), st)
        sk = max(1, min(16, R // 512))
        gemm_f32(dzo.t(), cat, G(self.lin.weight), accumulate=(sk == 1), split_k=sk, stream=st)
        call('dfol_colsum', ptr(dzo), dzo.stride(0), R, n_out, ptr(G(self.lin.bias)), st)
