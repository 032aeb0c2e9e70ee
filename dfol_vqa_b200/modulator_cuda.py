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

import os

import numpy as np
import torch

from . import capi
from .capi import call, ptr
from .engine import gemm_f32
from .modulator import OPS_NUM, AttentionTransfer


# One persistent kernel per pass (csrc/modulator_tape.cu) instead of one launch per cell; DFOL_MOD_TAPE=0 keeps the
# launch-per-step path (same arithmetic; the tests hold the two against each other)
_USE_TAPE = os.environ.get('DFOL_MOD_TAPE', '1') != '0'

_REC_DTYPE = np.dtype([('kind', np.int32), ('net', np.int32), ('rows', np.int32), ('base', np.int32),
                       ('live', np.int32), ('pad', np.int32), ('in_h', np.int64), ('in_c', np.int64),
                       ('add_h', np.int64), ('add_c', np.int64), ('fb_h', np.int64), ('fb_c', np.int64),
                       ('out_h', np.int64), ('out_c', np.int64), ('owner', np.uint64), ('mask', np.uint64),
                       ('part', np.uint64)])
_CELL, _OUT, _SQUEEZE, _GATE = 0, 1, 2, 3


def _carve(dev, sizes):
    """Zero-filled fp32 buffers of the given sizes out of ONE allocation / one memset (64-float aligned pieces)."""
    offs, total = [], 0
    for n in sizes:
        offs.append(total)
        total += (max(n, 1) + 63) // 64 * 64
    flat = torch.zeros(total, device=dev, dtype=torch.float32)
    return [flat[o:o + n] for o, n in zip(offs, sizes)]


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
        if _USE_TAPE:
            return self._forward_tape(cp)
        return self._forward_launches(cp)

    # ---- the walk over the op slots (shared by the launch-per-step path and the tape builder)

    @staticmethod
    def _walk(cp, plan, ops):
        """Forward pass over the slots, then the backward pass over the reversed dependencies, expressed with the
        primitive steps of ``ops``: cell(net, slot_key, state, add, owner, mask), out_layer(slot_key, fstate, bstate),
        squeeze(state, owner), gate(new, old, mask), empty() (the all-zero state)."""
        descs = cp.mod_descs
        n = len(descs)
        fwd_state = {}
        cell, out_layer, squeeze, gate, empty = ops.cell, ops.out_layer, ops.squeeze, ops.gate, ops.empty

        def run(i, d, is_forward, inputs, gated):
            """One slot of one pass; returns the outgoing state (or a pair for two-input terminals)."""
            op = d['op']
            mask = plan['masks'].get(i) if gated else None
            if op == 'select':
                if d['select'] is None:
                    return empty() if is_forward else inputs[0]
                if is_forward:
                    fwd_state[(i, 'select')] = cell('f', (i, 'select'), empty())
                    return fwd_state[(i, 'select')]
                out_layer((i, 'select'), fwd_state[(i, 'select')], inputs[0])
                return inputs[0]  # the select's own backward cell feeds nothing (first slot of its branch)
            if op in ('relate', 'verify_rel', 'choose_rel'):
                state = inputs[0]
                owner = plan['owners'].get((i, 'relate'))
                if is_forward:
                    x = empty()
                    if d['select'] is not None:
                        x = fwd_state[(i, 'select')] = cell('f', (i, 'select'), empty())
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
                inputs = [empty(), empty()] if two_in else [empty()]
            x = run(i, d, False, inputs, gated=(bool(d['deps']) and d['mask'] is not None and i != n - 1))
            back[i] = x

    def _forward_launches(self, cp):
        """One kernel launch per step (the path of round 1; DFOL_MOD_TAPE=0)."""
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
        lin = self.lin

        def zeros(rows):
            return torch.zeros(rows, S, device=dev, dtype=torch.float32)

        class Ops(object):
            @staticmethod
            def empty():
                return _State()

            @staticmethod
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
                     ptr(state.h), ptr(state.c), ptr(None if add is None else add.h),
                     ptr(None if add is None else add.c), ptr(owner), ptr(mask), ptr(None if fb is None else fb.h),
                     ptr(None if fb is None else fb.c), ptr(out.h), ptr(out.c), ptr(saved[net][base:]), rows, st)
                tape.append(('cell', net, base, rows, state, add if (add is not None and add.h is not None) else None,
                             owner, mask, fb, out))
                return out

            @staticmethod
            def out_layer(slot_key, fstate, bstate):
                base, rows = plan['base'][slot_key]
                call('dfol_mod_out_fwd', ptr(fstate.h), ptr(bstate.h), None, ptr(lin.weight), ptr(lin.bias), S,
                     n_out, ptr(mods[base:]), ptr(cat[base:]), rows, st)
                tape.append(('out', base, rows, fstate, bstate))

            @staticmethod
            def squeeze(state, owner):
                out = _State(zeros(B).index_add_(0, owner, state.h), zeros(B).index_add_(0, owner, state.c))
                tape.append(('squeeze', state, owner, out))
                return out

            @staticmethod
            def gate(new, old, mask):
                """mask ? new : old (rows), recorded for the backward pass."""
                if old.h is None:
                    old.h, old.c = zeros(new.h.shape[0]), zeros(new.h.shape[0])
                keep = (mask > 0).unsqueeze(1)
                out = _State(torch.where(keep, new.h, old.h), torch.where(keep, new.c, old.c))
                tape.append(('gate', new, old, keep, out))
                return out

        self._walk(cp, plan, Ops)
        ctx = {'tape': tape, 'saved': saved, 'cat': cat, 'mods': mods, 'plan': plan, 'R': R}
        return mods[:cp.mod_rows], ctx

    # ---- tape path: the steps compiled into records once per program batch, one persistent kernel per pass

    def _tape(self, cp, dev, plan):
        key = ('tape', str(dev))
        hit = cp.mod_cache.get(key)
        if hit is not None:
            return hit
        S, B = self.S, cp.question_num
        recs = []
        cursor = [0]

        class TS(object):   # a state of the pool: offset of h (c follows), or None = the all-zero state
            __slots__ = ('off', 'rows')

            def __init__(self):
                self.off, self.rows = None, 0

            def alloc(self, rows):
                self.off, self.rows = cursor[0], rows
                cursor[0] += (2 * rows * S + 3) // 4 * 4

            def h(self):
                return -1 if self.off is None else self.off

            def c(self):
                return -1 if self.off is None else self.off + self.rows * S

        def new_state(rows):
            t = TS()
            t.alloc(rows)
            return t

        def part_of(slot_key, rows):
            """row -> question map of a step of slot ``slot_key``: the (question-sorted) owner list of the slot's option
            predicates, or None when the step has one row per question."""
            i, k = slot_key
            own = plan['owners'].get((i, 'filter' if k.startswith('filter') else k))
            if own is None or rows != own.shape[0]:
                assert rows == B, (slot_key, rows, B)
                return None
            return own

        def rec(kind, net=0, rows=0, base=0, a=None, b=None, fb=None, out=None, owner=None, mask=None, part=None):
            none = TS()
            a, b, fb, out = a or none, b or none, fb or none, out or none
            recs.append({'kind': kind, 'net': net, 'rows': rows, 'base': base, 'a': a, 'b': b, 'fb': fb, 'out': out,
                         'owner': 0 if owner is None else owner.data_ptr(),
                         'mask': 0 if mask is None else mask.data_ptr(),
                         'part': 0 if part is None else part.data_ptr()})

        class Ops(object):
            @staticmethod
            def empty():
                return TS()

            @staticmethod
            def cell(net, slot_key, state, add=None, owner=None, mask=None):
                base, rows = plan['base'][slot_key]
                out = new_state(rows)
                fb = None
                if mask is not None:
                    fb = state
                    if fb.off is None:
                        fb.alloc(rows)   # the pool is zero-filled: a materialised zero state
                rec(_CELL, 0 if net == 'f' else 1, rows, base, a=state,
                    b=add if (add is not None and add.off is not None) else None, fb=fb, out=out, owner=owner, mask=mask,
                    part=part_of(slot_key, rows))
                return out

            @staticmethod
            def out_layer(slot_key, fstate, bstate):
                base, rows = plan['base'][slot_key]
                rec(_OUT, 0, rows, base, a=fstate, b=bstate, part=part_of(slot_key, rows))

            @staticmethod
            def squeeze(state, owner):
                out = new_state(B)
                assert owner.shape[0] == state.rows
                rec(_SQUEEZE, 0, state.rows, 0, a=state, out=out, owner=owner, part=owner)
                return out

            @staticmethod
            def gate(new, old, mask):
                if old.off is None:
                    old.alloc(new.rows)
                out = new_state(new.rows)
                assert new.rows == B
                rec(_GATE, 0, new.rows, 0, a=new, b=old, out=out, mask=mask)
                return out

        self._walk(cp, plan, Ops)
        # liveness of the backward pass: a cell / squeeze / gate whose output feeds nothing is skipped
        touched = set()
        for r in reversed(recs):
            k = r['kind']
            if k == _OUT:
                r['live'] = 1
                touched.add(id(r['a']))
                if r['b'].off is not None:
                    touched.add(id(r['b']))
                continue
            r['live'] = 1 if id(r['out']) in touched else 0
            if r['live']:
                for name in ('a', 'b', 'fb'):
                    if r[name].off is not None:
                        touched.add(id(r[name]))
        arr = np.zeros(len(recs), dtype=_REC_DTYPE)
        for i, r in enumerate(recs):
            arr[i] = (r['kind'], r['net'], r['rows'], r['base'], r['live'], 0, r['a'].h(), r['a'].c(), r['b'].h(),
                      r['b'].c(), r['fb'].h(), r['fb'].c(), r['out'].h(), r['out'].c(), r['owner'], r['mask'], r['part'])
        assert capi.lib().dfol_mod_tape_record_size() == _REC_DTYPE.itemsize
        dev_recs = torch.from_numpy(arr.view(np.uint8)).to(dev)
        hit = {'recs': dev_recs, 'n': len(recs), 'questions': B, 'pool': max(cursor[0], 4), 'host': arr}
        cp.mod_cache[key] = hit
        return hit

    def _forward_tape(self, cp):
        dev = self.fwd.weight_ih.device
        st = capi.stream_ptr(dev)
        S, n_out = self.S, self.n_out
        plan = self._plan(cp, dev)
        tp = self._tape(cp, dev, plan)
        F_all = plan['feats']
        R = F_all.shape[0]
        x_f = torch.empty(R, 4 * S, device=dev, dtype=torch.float32)
        x_b = torch.empty(R, 4 * S, device=dev, dtype=torch.float32)
        gemm_f32(F_all, self.fwd.weight_ih.t(), x_f, self.fwd.bias_ih, stream=st)
        gemm_f32(F_all, self.bwd.weight_ih.t(), x_b, self.bwd.bias_ih, stream=st)
        # one zero fill for everything the kernels accumulate into or may leave untouched
        pool, saved_f, saved_b, cat = _carve(dev, [tp['pool'], R * 7 * S, R * 7 * S, R * 2 * S])
        saved_f, saved_b, cat = saved_f.view(R, 7 * S), saved_b.view(R, 7 * S), cat.view(R, 2 * S)
        mods = torch.empty(R, n_out, device=dev, dtype=torch.float32)
        call('dfol_mod_tape_fwd', ptr(tp['recs']), tp['n'], tp['questions'], ptr(pool), ptr(x_f), ptr(x_b),
             ptr(self.fwd.weight_hh), ptr(self.fwd.bias_hh), ptr(self.bwd.weight_hh), ptr(self.bwd.bias_hh),
             ptr(self.lin.weight), ptr(self.lin.bias), S, n_out, ptr(saved_f), ptr(saved_b), ptr(mods), ptr(cat), st)
        ctx = {'tape_plan': tp, 'saved': {'f': saved_f, 'b': saved_b}, 'cat': cat, 'mods': mods, 'plan': plan, 'R': R,
               'keep': (pool, x_f, x_b)}
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
        tp = ctx.get('tape_plan')
        if tp is not None:
            gpool, dpf, dpb, dzo = _carve(dev, [tp['pool'], R * 4 * S, R * 4 * S, R * n_out])
            dpre = {'f': dpf.view(R, 4 * S), 'b': dpb.view(R, 4 * S)}
            dzo = dzo.view(R, n_out)
            call('dfol_mod_tape_bwd', ptr(tp['recs']), tp['n'], tp['questions'], ptr(gpool), ptr(self.fwd.weight_hh),
                 ptr(self.bwd.weight_hh), ptr(self.lin.weight), S, n_out, ptr(saved['f']), ptr(saved['b']), ptr(mods),
                 ptr(d_mods), ptr(dpre['f']), ptr(dpre['b']), ptr(dzo), st)
            return self._parameter_gradients(dpre, dzo, saved, cat, F_all, R, grads, st)
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

        return self._parameter_gradients(dpre, dzo, saved, cat, F_all, R, grads, st)

    def _parameter_gradients(self, dpre, dzo, saved, cat, F_all, R, grads, st):
        """dW_ih, dW_hh, the bias gradients and the output layer's: GEMM-shaped reductions over all cell rows."""
        S, n_out = self.S, self.n_out

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
