"""Tensor-core (bf16 operand, fp32 accumulate) scene build and backward of the reasoning path.

Every dense contraction of the visual oracle -- forward, dgrad and wgrad -- runs on tcgen05 (dfol_gemm_bf16_tc*),
the pair hidden layer and the table-layer backward run as single-pass HBM-bound kernels (csrc/tc_support.cu).
Reference being replaced: featurize_scene (nsvqa/data/batch_gqa_boxfeatures_pipeline.py:199-281),
ClassifierOracle.compute_all_log_likelihood_2 (nsvqa/nn/vision/classifier_oracle.py:145-156) and their autograd.

Weights stay fp32 (master copy, optimiser state); their bf16 operand copies -- plain and transposed, K padded to 64
-- are refreshed by ONE batched cast launch per step (dfol_cast_jobs).
"""

import os

import numpy as np
import torch

from . import capi
from .capi import K, call, ptr

DEFAULT_LL = -30.0

_REL_PTAB = os.environ.get('DFOL_REL_PTAB', '1') != '0'   # measurement switch
_JOB_DTYPE = np.dtype([('src', np.uint64), ('lds', np.int64), ('rows', np.int32), ('cols', np.int32),
                       ('dst', np.uint64), ('ldd', np.int64), ('out_rows', np.int32), ('transpose', np.int32),
                       ('dcols', np.int32), ('pad', np.int32)], align=True)


def _roundup(x, m):
    return (x + m - 1) // m * m


_FUSE_WGRAD = os.environ.get('DFOL_FUSE_WGRAD', '1') != '0'

class _Operands(object):
    """bf16 operand copies of the 12 parameter tensors + the device job table that refreshes them."""

    def __init__(self, w, device):
        assert capi.lib().dfol_cast_job_size() == _JOB_DTYPE.itemsize
        feat, a0, a1, r0, r1, emb = w.feat, w.attr[0], w.attr[1], w.rel[0], w.rel[1], w.emb
        F, D = feat.weight.shape
        ldo = F + 4
        Ha, H, E, C = a0.weight.shape[0], r0.weight.shape[0], r1.weight.shape[0], emb.weight.shape[0]
        assert a1.weight.shape[0] == E and emb.weight.shape[1] == E and r0.weight.shape[1] == 2 * ldo + 4
        self.dims = dict(F=F, D=D, ldo=ldo, Ha=Ha, H=H, E=E, C=C)
        Dp, Op, Hp, Hap, Ep = (_roundup(v, 64) for v in (D, ldo, H, Ha, E))
        Kc = Hap + 2 * Hp
        self.pad = dict(Dp=Dp, Op=Op, Hp=Hp, Hap=Hap, Ep=Ep, Kc=Kc)

        def buf(rows, cols):
            return torch.zeros(rows, cols, device=device, dtype=torch.bfloat16)

        self.wf = buf(F, Dp)
        self.wa1 = buf(Ha, Op)
        self.wa2 = buf(E, Hap)
        self.we = buf(C, Ep)
        self.wuv = buf(2 * H, Op)
        self.wr2 = buf(E, Hp)
        self.wa2t = buf(Ha, Ep)     # dgrad operand of attribute layer 2: [in, out]
        self.wr2t = buf(H, Ep)      # dgrad operand of relation layer 2
        self.wcat_t = buf(F, Kc)    # dgrad operand of the three first layers that read obj: [F, Ha | H | H]
        self.we_t = buf(Ep, _roundup(C, 64))  # dgrad operand of the table layer (dense gradients of query-type programs)
        jobs = []

        def job(src, rows, cols, dst, out_rows, dcols, transpose=0):
            assert src.stride(1) == 1 and dst.stride(1) == 1
            jobs.append((src.data_ptr(), src.stride(0), rows, cols, dst.data_ptr(), dst.stride(0), out_rows,
                         transpose, dcols, 0))

        fwd = [(feat.weight, self.wf), (a0.weight, self.wa1), (a1.weight, self.wa2), (emb.weight, self.we),
               (r1.weight, self.wr2)]
        for src, dst in fwd:
            job(src, src.shape[0], src.shape[1], dst, dst.shape[0], dst.shape[1])
        job(r0.weight[:, :ldo], H, ldo, self.wuv[:H], H, Op)
        job(r0.weight[:, ldo:2 * ldo], H, ldo, self.wuv[H:], H, Op)
        self.fwd_jobs = len(jobs)
        job(a1.weight, E, Ha, self.wa2t, Ha, Ep, 1)
        job(r1.weight, E, H, self.wr2t, H, Ep, 1)
        job(a0.weight[:, :F], Ha, F, self.wcat_t[:, :Hap], F, Hap, 1)
        job(r0.weight[:, :F], H, F, self.wcat_t[:, Hap:Hap + Hp], F, Hp, 1)
        job(r0.weight[:, ldo:ldo + F], H, F, self.wcat_t[:, Hap + Hp:], F, Hp, 1)
        job(emb.weight, C, E, self.we_t, E, self.we_t.shape[1], 1)
        arr = np.array(jobs, dtype=_JOB_DTYPE)
        self.max_elems = int(max(int(j[6]) * int(j[8]) for j in jobs))
        self.jobs = torch.from_numpy(arr.view(np.uint8).copy()).to(device)
        self.n_jobs = len(jobs)
        self.key = tuple(p.data_ptr() for p in w.parameters())

    def refresh(self, st, training):
        n = self.n_jobs if training else self.fwd_jobs
        call('dfol_cast_jobs', ptr(self.jobs), n, self.max_elems, st)


class TensorCorePath(object):

    def __init__(self, engine):
        self.engine = engine
        self.w = engine.w
        self._ops = None
        self._side = {}

    def side_stream(self, device):
        """Second stream for the attribute chain: its object-level GEMMs (T rows, < 148 tiles) leave most SMs idle, so
        they run beside the pair-level kernels of the relation chain; joined with events before the interpreter (forward)
        and before the shared first-layer gradients (backward)."""
        key = str(device)
        if key not in self._side:
            self._side[key] = torch.cuda.Stream(device)
        return self._side[key]

    def operands(self, device):
        key = tuple(p.data_ptr() for p in self.w.parameters())
        if self._ops is None or self._ops.key != key:
            w = self.w
            assert len(w.attr) == 2 and len(w.rel) == 2, 'bf16 path: one hidden layer per network (reference configs)'
            self._ops = _Operands(w, device)
        return self._ops

    @staticmethod
    def _slots_on_tensor_cores(cp):
        import os
        thr = int(os.environ.get('DFOL_SLOTS_TC_MIN', '1'))   # measured faster for every slot count (c1: 0.11 -> 0.07 ms)
        return cp.max_slots >= thr

    _P = -1   # pair rows of the scene being processed: kernel tags say "P" instead of the (batch-dependent) number

    @classmethod
    def _rows(cls, m):
        return 'P' if m == cls._P else '%d' % m

    @classmethod
    def _tc(cls, A16, B16, C, N, Kp, bias, act, st, table=None):
        """C = epilogue(A16[:, :Kp] @ B16[:N, :Kp]^T) on the tensor cores (dfol_gemm_bf16_tc)."""
        M = A16.shape[0]
        if table is None:
            out_bf16 = int(C.dtype == torch.bfloat16)
            ldc, store, maps, diag = C.stride(0), 0, (None, None, None, None, None), 0.0
        else:
            out_bf16, ldc, store = 0, 0, 1
            maps = (ptr(table['row_img']), ptr(table['img_row']), ptr(table['img_blk']), ptr(table['img_stride']),
                    ptr(table.get('img_n')))
            diag = table.get('diag', DEFAULT_LL)
        if capi.trace is not None:
            capi.next_meta = {'tag': 'gemm_bf16_tc[%sx%dx%d]%s' % (cls._rows(M), N, Kp, ' table' if store else ''),
                              'flops': 2.0 * M * N * Kp}
        call('dfol_gemm_bf16_tc', ptr(A16), A16.stride(0), ptr(B16), B16.stride(0), ptr(C), ldc, ptr(bias), M, N, Kp,
             act, out_bf16, store, maps[0], maps[1], maps[2], maps[3], maps[4], diag, st)

    @classmethod
    def _dgrad(cls, dZ, Wt, dX, N, Kp, h_saved, mul_mode, st, keep=1.0):
        """dX[:, :cols(dX)] = (dZ . Wt^T) * act'(h_saved); dX may be a column block of a wider buffer."""
        M = dZ.shape[0]
        if capi.trace is not None:
            capi.next_meta = {'tag': 'gemm_bf16_tc_dgrad[%sx%dx%d]' % (cls._rows(M), N, Kp), 'flops': 2.0 * M * N * Kp}
        call('dfol_gemm_bf16_tc_dgrad', ptr(dZ), dZ.stride(0), ptr(Wt), Wt.stride(0), ptr(dX), dX.stride(0), dX.shape[1], M,
             N, Kp, ptr(h_saved), 0 if h_saved is None else h_saved.stride(0), mul_mode, float(keep), st)

    @classmethod
    def _wgrad(cls, A, Mc, B, Nc, C, st):
        """C[Mc, Nc] (fp32 view) += A[:, :Mc]^T . B[:, :Nc] (reduction over rows)."""
        rows = A.shape[0]
        if capi.trace is not None:
            capi.next_meta = {'tag': 'gemm_bf16_tc_wgrad[%dx%dx%s]' % (Mc, Nc, cls._rows(rows)),
                              'flops': 2.0 * Mc * Nc * rows,
                              'bytes': 2.0 * rows * (A.stride(0) + B.stride(0)) if rows == cls._P else 0.0}
        call('dfol_gemm_bf16_tc_wgrad', ptr(A), A.stride(0), ptr(B), B.stride(0), ptr(C), C.stride(0), Mc, Nc, rows,
             st)

    @classmethod
    def _pair_dgrad_wgrad(cls, dz2, w2t, dz1, h1, gW2, P, H, Hp, E, Ep, keep, st):
        """Backward of the pair-level second layer: dZ1 = (dZ2 . W2) * elu'(H1) and dW2 += dZ2^T . H1.  One launch
        (``dfol_pair_layer_dgrad_wgrad_cluster``: both products from the operands the dgrad holds in shared memory, dZ2 and H1
        read from HBM once) when the layer fits its TMEM budget; ``DFOL_FUSE_WGRAD=0`` keeps the two launches."""
        fused = _FUSE_WGRAD and 128 < Hp <= 256 and 128 + Ep <= 512
        if not fused:
            cls._wgrad(dz2, E, h1, H, gW2, st)
        if capi.trace is not None:
            capi.next_meta = {'tag': 'pair_layer_dgrad%s_cluster[Px%dx%d]' % ('_wgrad' if fused else '', H, Ep),
                              'flops': (4.0 if fused else 2.0) * P * H * E, 'bytes': 2.0 * P * (Ep + 2 * Hp)}   # (algorithmic: E, not the padded Ep)
        if fused:
            call('dfol_pair_layer_dgrad_wgrad_cluster', ptr(dz2), Ep, ptr(w2t), Ep, ptr(dz1), Hp, Hp, P, H, Ep, ptr(h1),
                 Hp, K.MUL_ELU_GRAD, keep, ptr(gW2), gW2.stride(0), E, st)
        else:
            call('dfol_pair_layer_dgrad_cluster', ptr(dz2), Ep, ptr(w2t), Ep, ptr(dz1), Hp, Hp, P, H, Ep, ptr(h1), Hp,
                 K.MUL_ELU_GRAD, keep, st)

    # -------------------------------------------------------------------------------------------- forward

    def build_scene(self, features, layout, training, cp=None, dropout=None):
        from .engine import (Scene, DROP_FEATURES, DROP_ATTR_IN, DROP_ATTR_HIDDEN, DROP_REL_IN, DROP_REL_HIDDEN,
                             DROP_EMB_ATTR, DROP_EMB_REL)
        capi.lib()
        w = self.w
        staged = features if isinstance(features, tuple) else None   # (bf16 features (T, D), fp32 geometry (T, 6))
        dev = staged[1].device if staged is not None else features.device
        st = capi.stream_ptr(dev)
        ops = self.operands(dev)
        d, p = ops.dims, ops.pad
        F, D, ldo, Ha, H, E, C = d['F'], d['D'], d['ldo'], d['Ha'], d['H'], d['E'], d['C']
        if staged is not None:
            x16_in, geometry = staged
            T = geometry.shape[0]
            assert D == p['Dp'] and x16_in.shape == (T, D) and x16_in.dtype == torch.bfloat16 and x16_in.stride(1) == 1
            assert geometry.shape == (T, 6) and geometry.dtype == torch.float32 and geometry.is_contiguous()
            features = None
        else:
            T = features.shape[0]
            assert features.dtype == torch.float32 and features.stride(1) == 1 and features.shape[1] == D + 6
        assert T == layout.T
        TensorCorePath._P = layout.P
        ops.refresh(st, training)
        sc = Scene()
        sc.layout, sc.features, sc.tc, sc.dropout = layout, features, True, dropout

        def bf(rows, cols):
            return torch.empty(rows, cols, device=dev, dtype=torch.bfloat16)

        def drop(x, cols, site, stream=None):
            """x[:, :cols] *= mask of dropout site ``site`` (in place); no-op without dropout."""
            if dropout is not None:
                call('dfol_dropout_scale', ptr(x), x.stride(0), x.shape[0], cols, int(x.dtype == torch.bfloat16),
                     int(dropout[1]), site, float(dropout[0]), st if stream is None else stream)
            return x

        # featurizer: obj = [sigmoid(X Wf^T + b) | box position] (fp32 for the pair kernel) + bf16 operand copy
        if staged is not None:
            # the host already holds the features as bf16 (ProgramBatch.stage_bf16): same bits as the cast below
            x16 = x16_in if dropout is None else x16_in.clone()
            box, ldbox, box_col = geometry, 6, 0
        else:
            x16 = bf(T, p['Dp'])
            call('dfol_cast_bf16', ptr(features), features.stride(0), ptr(x16), p['Dp'], T, D, st)
            box, ldbox, box_col = features, features.stride(0), D
        drop(x16, D, DROP_FEATURES)
        obj = torch.empty(T, ldo, device=dev, dtype=torch.float32)
        self._tc(x16, ops.wf, obj, F, p['Dp'], w.feat.bias, K.ACT_SIGMOID, st)
        obj16 = bf(T, p['Op'])
        call('dfol_obj_finish', ptr(box), ldbox, box_col, ptr(obj), ldo, F, ptr(obj16), p['Op'], T, st)
        sc.obj, sc.obj16, sc.x16 = obj, obj16, x16

        # attribute chain (bf16 activations) -> attribute table (all C concept columns), on the side stream
        main = torch.cuda.current_stream(dev)
        # (per-kernel tracing keeps everything on one stream so that the CUDA-event table shows isolated kernel times)
        side = main if capi.trace is not None else self.side_stream(dev)
        ev_obj = torch.cuda.Event()
        ev_obj.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ev_obj)
            st2 = side.cuda_stream
            h1a = bf(T, p['Hap'])
            obj16a = obj16 if dropout is None else drop(obj16.clone(), ldo, DROP_ATTR_IN, st2)
            sc.obj16a = obj16a
            self._tc(obj16a, ops.wa1, h1a, Ha, p['Op'], w.attr[0].bias, K.ACT_ELU, st2)
            drop(h1a, Ha, DROP_ATTR_HIDDEN, st2)
            h2a = bf(T, p['Ep'])
            self._tc(h1a, ops.wa2, h2a, E, p['Hap'], w.attr[1].bias, K.ACT_SIGMOID, st2)
            drop(h2a, E, DROP_EMB_ATTR, st2)
            attr_ll = torch.empty(layout.attr_size, device=dev, dtype=torch.float32)
            obj_table = {'row_img': layout.obj_img, 'img_row': layout.obj_row, 'img_blk': layout.attr_blk,
                         'img_stride': layout.attr_stride}
            self._tc(h2a, ops.we, attr_ll, C, p['Ep'], w.emb.bias, K.ACT_LOGSIGMOID, st2, table=obj_table)
            ev_attr = torch.cuda.Event()
            ev_attr.record(side)
        sc.attr_ll = attr_ll
        sc.attr_h = [h1a, h2a]

        # relation chain: U|V in one GEMM, pair hidden layer, layer 2, relation table
        if layout.P == 0:
            # demand-driven pair rows: no program of this batch reads a relation likelihood -> no pair-level work at all
            assert cp is not None and cp.img_slot is not None
            dc = self.engine.upload_programs(cp, dev)
            sc.rel_ll = torch.empty(max(cp.rel_slot_size, 1), device=dev, dtype=torch.float32)
            sc.rel_blk, sc.rel_slots = dc['slot_blk'], True
            sc.rel_h, sc.uv = [None, None], None
            sc.geo = torch.empty(0, 4, device=dev, dtype=torch.float32) if training else None
            main.wait_event(ev_attr)
            return sc
        first = w.rel[0]
        h1r = bf(layout.P, p['Hp'])
        if dropout is None:
            uv = torch.empty(T, 2 * H, device=dev, dtype=torch.float32)
            self._tc(obj16, ops.wuv, uv, 2 * H, p['Op'], None, K.ACT_NONE, st)
            geo = torch.empty(layout.P, 4, device=dev, dtype=torch.float32) if training else None
            import os
            # one kernel for the hidden layer AND layer 2 (pair_chain_fwd.cu): the hidden layer is produced straight into
            # the shared-memory A operand of the layer-2 GEMM and never read back from HBM (bit-identical results)
            chain = (cp is not None and cp.img_slot is not None and H == 256 and p['Hp'] == 256 and 192 < p['Ep'] <= 384
                     and os.environ.get('DFOL_PAIR_CHAIN', '0') == '1'
                     and os.environ.get('DFOL_FUSED_INFERENCE', '0') != '1')
            if chain:
                if not training:
                    h1r = None
            elif 2 * layout.max_n + 5 <= 256 and H <= 256 and os.environ.get('DFOL_PAIR_HIDDEN_MMA', '0') == '1':
                # one-hot grouped GEMM on the tensor cores (pair_hidden_mma.cu).  Opt-in: correct, but its first version
                # (one tile per CTA, row-per-lane 32-byte stores) measures 0.181 ms against 0.163 ms for the SIMT kernel at
                # c1 and 1.13 against 0.63 ms at c3 (192 KB of shared memory: one CTA per SM, no overlap of its phases)
                if capi.trace is not None:
                    capi.next_meta = {'tag': 'pair_hidden_fwd_mma', 'bytes': 2.0 * layout.P * p['Hp']}
                Kp1 = _roundup(2 * layout.max_n + 5, 64)
                bm = bf(layout.B * H, Kp1)
                call('dfol_pair_hidden_fwd_mma', ptr(uv), uv.stride(0), ptr(obj[:, F:]), ldo,
                     ptr(first.weight[:, 2 * ldo:]), first.weight.stride(0), ptr(first.bias), ptr(h1r), p['Hp'], H,
                     ptr(geo), ptr(layout.pair_row), ptr(layout.obj_row), ptr(layout.img_np), layout.B, layout.max_n,
                     ptr(bm), st)
            else:
                if capi.trace is not None:
                    capi.next_meta = {'tag': 'pair_hidden_fwd_tc', 'bytes': 2.0 * layout.P * p['Hp']}
                call('dfol_pair_hidden_fwd_tc', ptr(uv), uv.stride(0), ptr(obj[:, F:]), ldo,
                     ptr(first.weight[:, 2 * ldo:]), first.weight.stride(0), ptr(first.bias), ptr(h1r), p['Hp'], H,
                     ptr(geo), ptr(layout.pair_row), ptr(layout.obj_row), ptr(layout.img_np), layout.B, layout.max_n,
                     st)
        else:
            chain = False
            # an independent mask per pair element breaks the U[s] + V[o] factorisation: the first layer runs as one
            # tcgen05 GEMM over the materialised, masked pair matrix (the reference's formulation,
            # batch_gqa_boxfeatures_pipeline.py:260-281)
            uv = geo = None
            width = 2 * ldo + 4
            Kp = _roundup(width, 64)
            pm = bf(layout.P, Kp)
            if capi.trace is not None:
                capi.next_meta = {'tag': 'pair_features_dropout', 'bytes': 2.0 * layout.P * Kp}
            call('dfol_pair_features_dropout', ptr(obj), ldo, ldo, F, ptr(pm), Kp, Kp, 1, ptr(layout.pair_row),
                 ptr(layout.obj_row), ptr(layout.img_n), ptr(layout.pair_img), layout.P, int(dropout[1]),
                 DROP_REL_IN, float(dropout[0]), st)
            w1 = torch.zeros(H, Kp, device=dev, dtype=torch.bfloat16)
            call('dfol_cast_bf16', ptr(first.weight), first.weight.stride(0), ptr(w1), Kp, H, width, st)
            if p['Hp'] > H:
                h1r.zero_()
            self._tc(pm, w1, h1r, H, Kp, first.bias, K.ACT_ELU, st)
            sc.pm, sc.w1_16 = (pm, w1) if training else (None, None)
            del pm
            drop(h1r, H, DROP_REL_HIDDEN)
        # the activation of layer 2 is only materialised when the backward pass (or the dense table) needs it
        slots = cp is not None and cp.img_slot is not None
        # inference with few relation columns per image: layer 2 and the columns in one kernel, no activation store
        # (opt-in: it saves the 377 MB H2 store of c1 but measures 0.38 ms against 0.19 + 0.11 ms for the cluster GEMM
        # followed by the slot columns -- the slot FMAs sit on its TMEM-drain path)
        import os
        fused = (slots and not training and cp.max_slots <= 4 and dropout is None
                 and os.environ.get('DFOL_FUSED_INFERENCE', '0') == '1')
        h2r = None if fused else bf(layout.P, p['Ep'])
        if slots:
            dc = self.engine.upload_programs(cp, dev)
            rel_ll = torch.empty(cp.rel_slot_size, device=dev, dtype=torch.float32)
            # probability table e^{ll} next to the log table (operand of the probability-space relate hop;
            # DFOL_REL_PTAB=0: the hop exponentiates the log tiles itself)
            rel_p = torch.empty(cp.rel_slot_size, device=dev, dtype=torch.float32) if _REL_PTAB else None
            if not fused:
                # the backward pass needs the layer-2 activation (or the image uses many relations): GEMM with bf16
                # store, then the demand-driven relation columns from the stored activation
                if chain:
                    if capi.trace is not None:
                        capi.next_meta = {'tag': 'pair_chain_fwd[Px%dx%d]' % (E, p['Hp']),
                                          'flops': 2.0 * layout.P * E * p['Hp'] + 10.0 * layout.P * H,
                                          'bytes': 2.0 * layout.P * ((p['Hp'] if training else 0) + p['Ep'])}
                    call('dfol_pair_chain_fwd', ptr(uv), uv.stride(0), ptr(obj[:, F:]), ldo,
                         ptr(first.weight[:, 2 * ldo:]), first.weight.stride(0), ptr(first.bias), ptr(ops.wr2), p['Hp'],
                         ptr(w.rel[1].bias), ptr(h2r), p['Ep'], p['Ep'], ptr(h1r), p['Hp'], ptr(geo),
                         ptr(layout.pair_img), ptr(layout.pair_row), ptr(layout.obj_row), ptr(layout.img_n), layout.P, E,
                         H, st)
                else:
                    if capi.trace is not None:
                        capi.next_meta = {'tag': 'pair_layer_fwd_cluster[Px%dx%d]' % (E, p['Hp']),
                                          'flops': 2.0 * layout.P * E * p['Hp'],
                                          'bytes': 2.0 * layout.P * (p['Hp'] + p['Ep'])}
                    call('dfol_pair_layer_fwd_cluster', ptr(h1r), p['Hp'], ptr(ops.wr2), p['Hp'], ptr(h2r), p['Ep'],
                         p['Ep'], ptr(w.rel[1].bias), layout.P, E, p['Hp'], K.ACT_SIGMOID, st)
                drop(h2r, E, DROP_EMB_REL)
                if capi.trace is not None:
                    capi.next_meta = {'tag': 'rel_slots_fwd', 'bytes': 2.0 * layout.P * p['Ep'] * max(
                        1, (cp.max_slots + 7) // 8) + 4.0 * cp.rel_slot_size}
                if self._slots_on_tensor_cores(cp):
                    # grouped tcgen05 GEMM: the activation is read once whatever the number of slots
                    if capi.trace is not None:
                        capi.next_meta = {'tag': 'rel_slots_fwd_tc', 'bytes': 2.0 * layout.P * p['Ep'] * max(
                            1, (cp.max_slots + 15) // 16) + 4.0 * cp.rel_slot_size}
                    wb = bf(16 * layout.B, p['Ep'])
                    call('dfol_rel_slots_fwd_tc', ptr(h2r), p['Ep'], layout.P, E, p['Ep'], ptr(w.emb.weight),
                         w.emb.weight.stride(0), ptr(w.emb.bias), ptr(dc['slot_wrow']), ptr(dc['img_slot']),
                         cp.max_slots, ptr(dc['slot_blk']), ptr(layout.rel_stride), ptr(layout.pair_row),
                         ptr(layout.img_nn), ptr(layout.img_np), layout.B, layout.max_n ** 2, DEFAULT_LL, ptr(wb),
                         ptr(rel_ll), ptr(rel_p), st)
                    sc.rel_p = rel_p
                else:
                    call('dfol_rel_slots_fwd', ptr(h2r), p['Ep'], E, ptr(w.emb.weight), w.emb.weight.stride(0),
                         ptr(w.emb.bias), ptr(dc['slot_wrow']), ptr(dc['img_slot']), cp.max_slots,
                         ptr(dc['slot_blk']), ptr(layout.rel_stride), ptr(layout.pair_row), ptr(layout.img_nn),
                         ptr(layout.img_np), layout.B, layout.max_n ** 2, DEFAULT_LL, ptr(rel_ll), ptr(rel_p), st)
                    sc.rel_p = rel_p
            else:
                # inference: layer 2 + relation columns in one persistent tcgen05 kernel; the P x E activation is
                # consumed in registers and never written
                if capi.trace is not None:
                    capi.next_meta = {'tag': 'pair_layer_fwd_tc[Px%dx%d]+slots' % (E, p['Hp']),
                                      'flops': 2.0 * layout.P * E * p['Hp']}
                call('dfol_pair_layer_fwd_tc', ptr(h1r), p['Hp'], ptr(ops.wr2), p['Hp'], None, p['Ep'], p['Ep'],
                     ptr(w.rel[1].bias), layout.P, E, p['Hp'], K.ACT_SIGMOID, ptr(w.emb.weight),
                     w.emb.weight.stride(0), ptr(w.emb.bias), ptr(dc['slot_wrow']), ptr(dc['img_slot']),
                     cp.max_slots, ptr(dc['slot_blk']), ptr(layout.rel_stride), ptr(layout.pair_img),
                     ptr(layout.pair_row), ptr(layout.img_n), DEFAULT_LL, ptr(rel_ll), st)
            sc.rel_blk, sc.rel_slots = dc['slot_blk'], True
        else:
            self._tc(h1r, ops.wr2, h2r, E, p['Hp'], w.rel[1].bias, K.ACT_SIGMOID, st)
            drop(h2r, E, DROP_EMB_REL)
            ridx = self.engine.rel_index(dev)
            sc.w_rel = w.emb.weight.detach().index_select(0, ridx).contiguous()
            sc.b_rel = w.emb.bias.detach().index_select(0, ridx).contiguous()
            wrel16 = bf(sc.w_rel.shape[0], p['Ep'])
            call('dfol_cast_bf16', ptr(sc.w_rel), sc.w_rel.stride(0), ptr(wrel16), p['Ep'], sc.w_rel.shape[0], E, st)
            rel_ll = torch.empty(layout.rel_size, device=dev, dtype=torch.float32)
            pair_table = {'row_img': layout.pair_img, 'img_row': layout.pair_row, 'img_blk': layout.rel_blk,
                          'img_stride': layout.rel_stride, 'img_n': layout.img_n, 'diag': DEFAULT_LL}
            self._tc(h2r, wrel16, rel_ll, sc.w_rel.shape[0], p['Ep'], sc.b_rel, K.ACT_LOGSIGMOID, st,
                     table=pair_table)
            sc.rel_blk, sc.rel_slots = layout.rel_blk, False
        sc.rel_ll = rel_ll
        sc.rel_h = [h1r, h2r]
        sc.uv, sc.geo = uv, geo
        main.wait_event(ev_attr)  # the interpreter reads both tables
        return sc

    # -------------------------------------------------------------------------------------------- backward

    def _table_backward(self, g, tabs, ll, blk, stride, row0, img_rows, max_rows, rows_total, W, dW, db, h_last,
                        d_below, st, tag, keep=1.0, remask=None, tile_tab=None):
        """dZ (bf16, rows x padded width) of the layer below a table layer, from the compact gradient slices."""
        dev = ll.device
        E = W.shape[1]
        cols = h_last.shape[1]
        dz = torch.empty(rows_total, cols, device=dev, dtype=torch.bfloat16)
        # (measured: the SIMT kernel's cost grows with the slices per image -- c3, 9-12 slices: 1.75 ms against 1.01 ms;
        # c2 / c4: 0.60 against 0.37 ms -- while with the 1-3 slices of c1 it is level: 0.17 against 0.18 ms)
        mma_min = int(os.environ.get('DFOL_TBL_MMA_MIN', '4'))
        if (tile_tab is not None and mma_min <= tabs['max_per_image'] <= 32 and keep == 1.0 and 128 <= cols <= 320
                and cols % 64 == 0):
            # tcgen05 version (table_layer_bwd_mma.cu): the three contractions of the tile as MMAs, H read once
            if capi.trace is not None:
                S = tabs['max_per_image']
                capi.next_meta = {'tag': 'table_layer_bwd_mma[%s]' % tag, 'bytes': 4.0 * rows_total * cols,
                                  'flops': 2.0 * rows_total * cols * (2 * (16 if S <= 16 else 32))}   # MMA-A + MMA-B
            wb = torch.empty(len(tabs['img_slice']) - 1, 32 * cols, device=dev, dtype=torch.bfloat16)
            call('dfol_table_layer_bwd_mma', ptr(g), ptr(tabs['goff']), ptr(tabs['col']), ptr(tabs['wrow']),
                 ptr(tabs['img_slice']), len(tabs['img_slice']) - 1, tabs['max_per_image'], ptr(ll), ptr(blk),
                 ptr(stride), ptr(row0), ptr(img_rows), ptr(tile_tab[0]), tile_tab[1], rows_total, ptr(W), W.stride(0),
                 ptr(h_last), h_last.stride(0), E, ptr(dz), dz.stride(0), cols, ptr(dW), ptr(db), ptr(d_below), ptr(wb),
                 st)
            return dz
        if tabs['max_per_image'] <= 24:
            if capi.trace is not None:
                passes = max(1, (tabs['max_per_image'] + 11) // 12)
                capi.next_meta = {'tag': 'table_layer_bwd_tc[%s]' % tag, 'bytes': 4.0 * rows_total * cols * passes}
            call('dfol_table_layer_bwd_tc', ptr(g), ptr(tabs['goff']), ptr(tabs['col']), ptr(tabs['wrow']),
                 ptr(tabs['img_slice']), len(tabs['img_slice']) - 1, max_rows, tabs['max_per_image'], ptr(ll),
                 ptr(blk), ptr(stride), ptr(row0), ptr(img_rows), ptr(W), W.stride(0), ptr(h_last),
                 h_last.stride(0), E, ptr(dz), dz.stride(0), cols, ptr(dW), ptr(db), ptr(d_below), float(keep), st)
            return dz
        assert keep == 1.0 or rows_total < (1 << 22), 'dropout: the dense table backward is meant for object rows'
        # many columns per image (option lists of query-type programs): dense d logits, then plain GEMMs
        from .engine import gemm_f32, _split_for
        Cn = dW.shape[0]
        dl = torch.zeros(rows_total, Cn, device=dev, dtype=torch.float32)
        call('dfol_table_grad_dense', ptr(g), ptr(tabs['goff']), ptr(tabs['col']), ptr(tabs['img']), tabs['count'],
             ptr(ll), ptr(blk), ptr(stride), ptr(row0), ptr(img_rows), ptr(dl), Cn, st)
        call('dfol_colsum', ptr(dl), Cn, rows_total, Cn, ptr(db), st)
        if keep == 1.0 and W is self.w.emb.weight and os.environ.get('DFOL_DENSE_FP32', '0') != '1':
            # tensor cores: d logits as a bf16 operand (like every other gradient operand of this mode);
            # dW += dl^T . h (MN-major wgrad), dZ = (dl . W) * sigmoid'(h) in the dgrad epilogue
            ops = self.operands(dev)
            Cp = ops.we_t.shape[1]
            dl16 = torch.empty(rows_total, Cp, device=dev, dtype=torch.bfloat16)
            call('dfol_cast_bf16', ptr(dl), Cn, ptr(dl16), Cp, rows_total, Cn, st)
            del dl
            self._wgrad(dl16, Cn, h_last, E, dW, st)
            # (N = the padded width: rows E.. of we_t are zero and sigmoid' of the zero padding of h is zero)
            self._dgrad(dl16, ops.we_t, dz, cols, Cp, h_last, K.MUL_SIGMOID_GRAD, st)
            call('dfol_colsum_bf16', ptr(dz), cols, rows_total, E, ptr(d_below), st)
            return dz
        h32 = h_last[:, :E].float()
        sk = _split_for(rows_total)
        gemm_f32(dl.t(), h32, dW, accumulate=(sk == 1), split_k=sk, stream=st)
        d = torch.empty(rows_total, E, device=dev, dtype=torch.float32)
        gemm_f32(dl, W, d, stream=st)
        if keep < 1.0:  # h_last is the post-dropout activation: re-mask the gradient, sigmoid' at h = saved * keep
            remask(d, E)
            h32 = h32 * keep
        call('dfol_act_grad_mul', ptr(d), E, ptr(h32), E, rows_total, E, K.ACT_SIGMOID, st)
        call('dfol_colsum', ptr(d), E, rows_total, E, ptr(d_below), st)
        call('dfol_cast_bf16', ptr(d), E, ptr(dz), cols, rows_total, E, st)
        return dz

    def backward(self, cp, scene, tape, d_lp, grads, early_hook=None):
        eng = self.engine
        w = self.w
        lay = scene.layout
        dev = scene.attr_ll.device
        st = capi.stream_ptr(dev)
        ops = self.operands(dev)
        d, p = ops.dims, ops.pad
        F, D, ldo, Ha, H, E = d['F'], d['D'], d['ldo'], d['Ha'], d['H'], d['E']
        Hap, Hp, Ep, Kc = p['Hap'], p['Hp'], p['Ep'], p['Kc']
        T, P = lay.T, lay.P
        TensorCorePath._P = P
        if getattr(scene, 'dropout', None) is not None:
            return self._backward_dropout(cp, scene, tape, d_lp, grads)   # (no early hook: one all-reduce at the end)
        assert scene.geo is not None, 'scene was built without training buffers'

        def G(prm):
            return grads[id(prm)]

        g_attr, g_rel = eng.program_backward(cp, scene, tape, d_lp)

        # dcat = [dZ1 of the attribute chain | dU | dV]: the three first layers that read obj share one dgrad and
        # (when their widths are equal multiples of 128) one segmented wgrad
        dcat = torch.zeros(T, Kc, device=dev, dtype=torch.bfloat16)
        merged = Ha == Hap == Hp == H and Hp % 128 == 0
        a0, a1, r0, r1 = w.attr[0], w.attr[1], w.rel[0], w.rel[1]

        # ---- attribute table layer -> layer 2 -> layer 1 (side stream, beside the relation chain)
        main = torch.cuda.current_stream(dev)
        side = main if capi.trace is not None else self.side_stream(dev)
        sa = eng._slice_tables(cp.attr_slices, lay.B, dev, 'attr_slices', cp)  # (first use uploads on `main`)
        ev_pb = torch.cuda.Event()
        ev_pb.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ev_pb)
            st2 = side.cuda_stream
            if sa['count']:
                dz2a = self._table_backward(g_attr, sa, scene.attr_ll, lay.attr_blk, lay.attr_stride, lay.obj_row,
                                            lay.img_n, lay.max_n, T, w.emb.weight, G(w.emb.weight), G(w.emb.bias),
                                            scene.attr_h[1], G(a1.bias), st2, 'attr')
                self._wgrad(dz2a, E, scene.attr_h[0], Ha, G(a1.weight), st2)
                self._dgrad(dz2a, ops.wa2t, dcat[:, :Hap], Ha, Ep, scene.attr_h[0], K.MUL_ELU_GRAD, st2)
                call('dfol_colsum_bf16', ptr(dcat), Kc, T, Ha, ptr(G(a0.bias)), st2)
                if not merged:
                    self._wgrad(dcat, Ha, scene.obj16, ldo, G(a0.weight), st2)
            ev_attr = torch.cuda.Event()
            ev_attr.record(side)

        # ---- relation table layer -> layer 2 -> pair hidden layer
        sr = eng._slice_tables(cp.rel_slices, lay.B, dev, 'rel_slices', cp)
        if sr['count']:
            if scene.rel_slots:
                dz2r = self._table_backward(g_rel, sr, scene.rel_ll, scene.rel_blk, lay.rel_stride, lay.pair_row,
                                            lay.img_nn, lay.max_n ** 2, P, w.emb.weight, G(w.emb.weight),
                                            G(w.emb.bias), scene.rel_h[1], G(r1.bias), st, 'rel',
                                            tile_tab=(lay.pair_tile, lay.pair_tiles))
            else:
                nR = scene.w_rel.shape[0]
                dw_rel = torch.zeros(nR, E, device=dev, dtype=torch.float32)
                db_rel = torch.zeros(nR, device=dev, dtype=torch.float32)
                dz2r = self._table_backward(g_rel, sr, scene.rel_ll, lay.rel_blk, lay.rel_stride, lay.pair_row,
                                            lay.img_nn, lay.max_n ** 2, P, scene.w_rel, dw_rel, db_rel,
                                            scene.rel_h[1], G(r1.bias), st, 'rel')
                ridx = eng.rel_index(dev)
                G(w.emb.weight).index_add_(0, ridx, dw_rel)
                G(w.emb.bias).index_add_(0, ridx, db_rel)
            h1r = scene.rel_h[0]
            dz1r = torch.empty(P, Hp, device=dev, dtype=torch.bfloat16)
            self._pair_dgrad_wgrad(dz2r, ops.wr2t, dz1r, h1r, G(r1.weight), P, H, Hp, E, Ep, 1.0, st)
            gw1 = G(r0.weight)
            if capi.trace is not None:
                capi.next_meta = {'tag': 'pair_hidden_bwd_tc', 'bytes': 2.0 * P * H + 16.0 * P}
            call('dfol_pair_hidden_bwd_tc', ptr(dz1r), Hp, ptr(scene.geo), ptr(dcat[:, Hap:]), ptr(dcat[:, Hap + Hp:]),
                 Kc, ptr(gw1[:, 2 * ldo:]), gw1.stride(0), ptr(G(r0.bias)), H, ptr(lay.pair_row), ptr(lay.obj_row),
                 ptr(lay.img_np), lay.B, lay.max_n, st)
            if not merged:
                self._wgrad(dcat[:, Hap:], H, scene.obj16, ldo, gw1[:, :ldo], st)
                self._wgrad(dcat[:, Hap + Hp:], H, scene.obj16, ldo, gw1[:, ldo:2 * ldo], st)

        main.wait_event(ev_attr)  # dcat[:, :Hap] and the attribute-side gradients are complete
        if early_hook is not None:
            # table layers, second layers, pair hidden layer and all biases but the featurizer's are done: their share
            # of the gradient bucket can be all-reduced while the first-layer / featurizer kernels below run
            early_hook()
        if merged:
            gw1 = G(r0.weight)
            if capi.trace is not None:
                capi.next_meta = {'tag': 'gemm_bf16_tc_wgrad_seg[%dx%dx%d]' % (Kc, ldo, T), 'flops': 2.0 * Kc * ldo * T}
            call('dfol_gemm_bf16_tc_wgrad_seg', ptr(dcat), Kc, ptr(scene.obj16), scene.obj16.stride(0),
                 ptr(G(a0.weight)), G(a0.weight).stride(0), ptr(gw1[:, :ldo]), gw1.stride(0),
                 ptr(gw1[:, ldo:2 * ldo]), gw1.stride(0), Hp, Kc, ldo, T, st)

        # ---- featurizer: d pre = (dcat . [Wa1 | Wu | Wv][:, :F]) * f (1 - f)
        dpre = torch.empty(T, F, device=dev, dtype=torch.bfloat16)
        self._dgrad(dcat, ops.wcat_t, dpre, F, Kc, scene.obj16, K.MUL_SIGMOID_GRAD, st)
        call('dfol_colsum_bf16', ptr(dpre), F, T, F, ptr(G(w.feat.bias)), st)
        self._wgrad(dpre, F, scene.x16, D, G(w.feat.weight), st)

    def _backward_dropout(self, cp, scene, tape, d_lp, grads):
        """Backward pass of a scene built with training-mode dropout.  Every saved activation is the tensor AFTER its
        dropout mask (the real operand of the next GEMM, hence of its wgrad); the kernels that need act'(h) recover h
        and the mask from it (``keep`` = 1 - p: h = saved * keep, mask factor (saved != 0) / keep), the masks of the
        two network inputs are recomputed (dfol_dropout_scale on the gradient).  The first relation layer runs on the
        materialised masked pair matrix, so its weight gradient is one tcgen05 wgrad and its input gradient one dgrad +
        mask + reduction over pairs (dfol_pair_features_bwd) instead of the factored pair-hidden backward."""
        from .engine import DROP_ATTR_IN, DROP_REL_IN, DROP_EMB_ATTR
        eng = self.engine
        w = self.w
        lay = scene.layout
        dev = scene.attr_ll.device
        st = capi.stream_ptr(dev)
        ops = self.operands(dev)
        d, p = ops.dims, ops.pad
        F, D, ldo, Ha, H, E = d['F'], d['D'], d['ldo'], d['Ha'], d['H'], d['E']
        Hap, Hp, Ep, Op = p['Hap'], p['Hp'], p['Ep'], p['Op']
        T, P = lay.T, lay.P
        prob, seed = float(scene.dropout[0]), int(scene.dropout[1])
        keep = 1.0 - prob
        assert scene.pm is not None, 'scene was built without training buffers'

        def G(prm):
            return grads[id(prm)]

        def remask(x, cols, site):
            call('dfol_dropout_scale', ptr(x), x.stride(0), x.shape[0], cols, int(x.dtype == torch.bfloat16), seed, site,
                 prob, st)

        a0, a1, r0, r1 = w.attr[0], w.attr[1], w.rel[0], w.rel[1]
        g_attr, g_rel = eng.program_backward(cp, scene, tape, d_lp)

        # ---- attribute chain
        d_obj_a = torch.zeros(T, Op, device=dev, dtype=torch.bfloat16)   # its share of d obj (masked input gradient)
        sa = eng._slice_tables(cp.attr_slices, lay.B, dev, 'attr_slices', cp)
        if sa['count']:
            h1a, h2a = scene.attr_h
            dz2a = self._table_backward(g_attr, sa, scene.attr_ll, lay.attr_blk, lay.attr_stride, lay.obj_row,
                                        lay.img_n, lay.max_n, T, w.emb.weight, G(w.emb.weight), G(w.emb.bias), h2a,
                                        G(a1.bias), st, 'attr', keep, lambda x, c: remask(x, c, DROP_EMB_ATTR))
            self._wgrad(dz2a, E, h1a, Ha, G(a1.weight), st)
            dz1a = torch.zeros(T, Hap, device=dev, dtype=torch.bfloat16)
            self._dgrad(dz2a, ops.wa2t, dz1a, Ha, Ep, h1a, K.MUL_ELU_GRAD, st, keep)
            call('dfol_colsum_bf16', ptr(dz1a), Hap, T, Ha, ptr(G(a0.bias)), st)
            self._wgrad(dz1a, Ha, scene.obj16a, ldo, G(a0.weight), st)
            self._dgrad(dz1a, ops.wcat_t[:, :Hap], d_obj_a, F, Hap, None, K.MUL_NONE, st)
            remask(d_obj_a, ldo, DROP_ATTR_IN)

        # ---- relation chain
        d_obj = torch.zeros(T, ldo, device=dev, dtype=torch.float32)
        sr = eng._slice_tables(cp.rel_slices, lay.B, dev, 'rel_slices', cp)
        if sr['count']:
            assert scene.rel_slots, 'dropout with trainable oracle networks needs the demand-driven relation table'
            h1r, h2r = scene.rel_h
            dz2r = self._table_backward(g_rel, sr, scene.rel_ll, scene.rel_blk, lay.rel_stride, lay.pair_row,
                                        lay.img_nn, lay.max_n ** 2, P, w.emb.weight, G(w.emb.weight), G(w.emb.bias),
                                        h2r, G(r1.bias), st, 'rel', keep)
            dz1r = torch.empty(P, Hp, device=dev, dtype=torch.bfloat16)
            self._pair_dgrad_wgrad(dz2r, ops.wr2t, dz1r, h1r, G(r1.weight), P, H, Hp, E, Ep, keep, st)
            del dz2r
            call('dfol_colsum_bf16', ptr(dz1r), Hp, P, H, ptr(G(r0.bias)), st)
            pm, w1 = scene.pm, scene.w1_16
            width = 2 * ldo + 4
            Kp = pm.shape[1]
            self._wgrad(dz1r, H, pm, width, G(r0.weight), st)
            # d pm = dZ1 . W1 (all pair columns), re-masked, reduced over pairs back to the objects
            w1t = w1.t().contiguous()   # [Kp, H]: dgrad operand (in x out)
            if Hp > H:
                w1t = torch.nn.functional.pad(w1t, (0, Hp - H))
            d_pm = torch.empty(P, Kp, device=dev, dtype=torch.bfloat16)
            self._dgrad(dz1r, w1t, d_pm, Kp, Hp, None, K.MUL_NONE, st)
            remask(d_pm, width, DROP_REL_IN)
            call('dfol_pair_features_bwd', ptr(d_pm), Kp, 1, ldo, ptr(d_obj), ldo, ptr(d_obj_a), Op,
                 ptr(lay.pair_row), ptr(lay.obj_row), ptr(lay.img_n), ptr(lay.obj_img), T, st)
        else:
            d_obj[:, :F] = d_obj_a[:, :F].float()

        # ---- featurizer: d pre = d obj[:, :F] * f (1 - f) (obj itself is not masked: the masks sit on its copies)
        call('dfol_act_grad_mul', ptr(d_obj), ldo, ptr(scene.obj), ldo, T, F, K.ACT_SIGMOID, st)
        call('dfol_colsum', ptr(d_obj), ldo, T, F, ptr(G(w.feat.bias)), st)
        dpre = torch.empty(T, F, device=dev, dtype=torch.bfloat16)
        call('dfol_cast_bf16', ptr(d_obj), ldo, ptr(dpre), F, T, F, st)
        self._wgrad(dpre, F, scene.x16, D, G(w.feat.weight), st)
