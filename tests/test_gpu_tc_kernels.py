"""GPU tests of the tensor-core-mode support kernels (csrc/tc_support.cu, program_*_fast.cu) through the C ABI:
operand preparation, pair hidden layer forward/backward, table-layer backward, the fast interpreter builds against the
exact ones, the demand-driven relation slots against the dense table, and the host staging pipeline."""

import json

import numpy as np
import pytest
import torch

import helpers
import dfol_oracle as orc  # noqa: F401  (path setup through helpers)

pytestmark = pytest.mark.gpu


def _geo(pos_s, pos_o):
    """Reference pair geometry (batch_gqa_boxfeatures_pipeline.py:260-279) for position rows [x, y, w, h]."""
    dx = pos_s[:, 0] + pos_s[:, 2] / 2 - pos_o[:, 0] - pos_o[:, 2] / 2
    dy = pos_s[:, 1] + pos_s[:, 3] / 2 - pos_o[:, 1] - pos_o[:, 3] / 2
    dist = torch.sqrt(dx * dx + dy * dy)
    ang = torch.asin(dy / dist.clamp(min=1e-10))
    return torch.stack([dist, ang, torch.sign(pos_o[:, 0] - pos_s[:, 0]), torch.sign(pos_o[:, 1] - pos_s[:, 1])], 1)


def _pair_index(counts):
    t0, rows = 0, []
    for n in counts:
        s = torch.arange(n).repeat_interleave(n) + t0
        o = torch.arange(n).repeat(n) + t0
        rows.append(torch.stack([s, o], 1))
        t0 += n
    return torch.cat(rows)


def _layout(counts):
    n = torch.tensor(counts)
    dev = lambda t, d: t.to(d).cuda()
    return {'img_n': dev(n, torch.int32), 'img_nn': dev(n * n, torch.int32),
            'obj_row': dev(torch.cat([torch.zeros(1, dtype=torch.long), n.cumsum(0)]), torch.int32),
            'pair_row': dev(torch.cat([torch.zeros(1, dtype=torch.long), (n * n).cumsum(0)]), torch.int32)}


@pytest.mark.parametrize('counts', [[48] * 5, [5, 12, 3, 33, 16, 1], [100, 64], [20, 31, 7, 32], [64, 50, 2], [128, 3]])
def test_pair_hidden_fwd_bwd_tc(counts):
    from dfol_vqa_b200.capi import call, ptr, stream_ptr
    g = torch.Generator().manual_seed(len(counts) + counts[0])
    H, T, P = 256, sum(counts), sum(c * c for c in counts)
    lay = _layout(counts)
    uv = (torch.randn(T, 2 * H, generator=g) * 0.5).cuda()
    obj = torch.zeros(T, 8)
    obj[:, 4:] = torch.rand(T, 4, generator=g)
    obj = obj.cuda()
    wfull = (torch.randn(H, 12, generator=g) * 0.3).cuda()   # geometry weights are the last four columns
    bias = (torch.randn(H, generator=g) * 0.1).cuda()
    h = torch.full((P, H), float('nan'), device='cuda', dtype=torch.bfloat16)
    geo = torch.full((P, 4), float('nan'), device='cuda')
    call('dfol_pair_hidden_fwd_tc', ptr(uv), 2 * H, ptr(obj[:, 4:]), 8, ptr(wfull[:, 8:]), 12, ptr(bias), ptr(h), H, H,
         ptr(geo), ptr(lay['pair_row']), ptr(lay['obj_row']), ptr(lay['img_n']), len(counts), max(counts), stream_ptr())
    torch.cuda.synchronize()
    idx = _pair_index(counts).cuda()
    s, o = idx[:, 0], idx[:, 1]
    gref = _geo(obj[s, 4:].double(), obj[o, 4:].double())
    offdiag = (s != o)
    # asin is ill-conditioned next to +-pi/2 (fp32 kernel vs fp64 here)
    assert torch.allclose(geo[offdiag].double(), gref[offdiag], rtol=1e-3, atol=5e-3)
    assert bool((geo[~offdiag] == 0).all())
    gz = torch.where(offdiag[:, None], gref, torch.zeros_like(gref))
    z = uv[s, :H].double() + uv[o, H:].double() + gz @ wfull[:, 8:].double().t() + bias.double()
    ref = torch.nn.functional.elu(z)
    assert torch.allclose(h.double(), ref, rtol=1.2e-2, atol=1.2e-2), (h.double() - ref).abs().max()

    # backward from a random bf16 dZ (rows of self pairs must be ignored)
    dz = (torch.randn(P, H, generator=g) * 0.1).cuda().bfloat16()
    dcat = torch.full((T, 2 * H), float('nan'), device='cuda', dtype=torch.bfloat16)
    dwg = torch.zeros(H, 12, device='cuda')
    db = torch.zeros(H, device='cuda')
    call('dfol_pair_hidden_bwd_tc', ptr(dz), H, ptr(geo), ptr(dcat), ptr(dcat[:, H:]), 2 * H, ptr(dwg[:, 8:]), 12,
         ptr(db), H, ptr(lay['pair_row']), ptr(lay['obj_row']), ptr(lay['img_n']), len(counts), max(counts),
         stream_ptr())
    torch.cuda.synchronize()
    d = torch.where(offdiag[:, None], dz.double(), torch.zeros(1, device='cuda', dtype=torch.float64))
    du = torch.zeros(T, H, device='cuda', dtype=torch.float64).index_add_(0, s, d)
    dv = torch.zeros(T, H, device='cuda', dtype=torch.float64).index_add_(0, o, d)
    scale = float(du.abs().max())
    assert float((dcat[:, :H].double() - du).abs().max()) <= 1e-2 * scale + 1e-3
    assert float((dcat[:, H:].double() - dv).abs().max()) <= 1e-2 * scale + 1e-3
    assert torch.allclose(db.double(), d.sum(0), rtol=1e-3, atol=1e-3)
    assert torch.allclose(dwg[:, 8:].double(), d.t() @ gz, rtol=1e-3, atol=1e-3)
    assert bool((dwg[:, :8] == 0).all())


@pytest.mark.parametrize('counts,slices_per_image', [([6, 10, 4], [1, 0, 3]), ([10, 8, 12, 5], [9, 2, 4, 5])])
def test_table_layer_bwd_tc(counts, slices_per_image):
    """dZ / dW / db / dbelow of the table layer from compact gradient slices vs an fp64 evaluation."""
    from dfol_vqa_b200.capi import call, ptr, stream_ptr
    g = torch.Generator().manual_seed(sum(counts))
    E, C, ld = 300, 40, 320
    rows = [c * c for c in counts]
    R = sum(rows)
    stride = [(r + 3) // 4 * 4 for r in rows]
    H2 = torch.zeros(R, ld, dtype=torch.bfloat16)
    H2[:, :E] = torch.rand(R, E, generator=g).bfloat16()
    W = (torch.randn(C, E, generator=g) * 0.3)
    blk, goff, cols, wrow, img_slice, total_g = [0], [], [], [], [0], 0
    for b, k in enumerate(slices_per_image):
        blk.append(blk[-1] + k * stride[b])
        for j in range(k):
            goff.append(total_g)
            cols.append(j)
            wrow.append(int(torch.randint(0, C, (1,), generator=g)))
            total_g += stride[b]
        img_slice.append(len(goff))
    ll = -torch.rand(max(blk[-1], 1), generator=g) * 3
    gsl = torch.randn(max(total_g, 1), generator=g) * (torch.rand(max(total_g, 1), generator=g) < 0.7)
    row0 = [0]
    for r in rows:
        row0.append(row0[-1] + r)
    i32 = lambda a: torch.tensor(a if len(a) else [0], dtype=torch.int32).cuda()
    dZ = torch.full((R, ld), float('nan'), dtype=torch.bfloat16, device='cuda')
    dW = torch.zeros(C, E, device='cuda')
    db = torch.zeros(C, device='cuda')
    dbelow = torch.zeros(E, device='cuda')
    # device copies must stay referenced until the launch has run (ptr() only takes the address)
    d = dict(g=gsl.cuda(), goff=i32(goff), cols=i32(cols), wrow=i32(wrow), img_slice=i32(img_slice), ll=ll.cuda(),
             blk=torch.tensor(blk[:-1], dtype=torch.int64).cuda(), stride=i32(stride), row0=i32(row0[:-1]),
             rows=i32(rows), W=W.cuda(), H2=H2.cuda())
    call('dfol_table_layer_bwd_tc', ptr(d['g']), ptr(d['goff']), ptr(d['cols']), ptr(d['wrow']), ptr(d['img_slice']),
         len(counts), max(rows), max(slices_per_image), ptr(d['ll']), ptr(d['blk']), ptr(d['stride']), ptr(d['row0']),
         ptr(d['rows']), ptr(d['W']), E, ptr(d['H2']), ld, E, ptr(dZ), ld, ld, ptr(dW), ptr(db), ptr(dbelow), 1.0,
         stream_ptr())
    torch.cuda.synchronize()
    h = H2[:, :E].double()
    ref_dz = torch.zeros(R, E, dtype=torch.float64)
    ref_dw = torch.zeros(C, E, dtype=torch.float64)
    ref_db = torch.zeros(C, dtype=torch.float64)
    for b, k in enumerate(slices_per_image):
        for j in range(k):
            sidx = img_slice[b] + j
            gl = gsl[goff[sidx]:goff[sidx] + rows[b]].double()
            l = ll[blk[b] + j * stride[b]: blk[b] + j * stride[b] + rows[b]].double()
            dz = gl * (1 - l.exp())
            hb = h[row0[b]:row0[b + 1]]
            ref_dz[row0[b]:row0[b + 1]] += dz[:, None] * W[wrow[sidx]].double()[None, :]
            ref_dw[wrow[sidx]] += dz @ hb
            ref_db[wrow[sidx]] += dz.sum()
    ref_dz = ref_dz * h * (1 - h)
    scale = float(ref_dz.abs().max()) + 1e-6
    assert float((dZ[:, :E].double().cpu() - ref_dz).abs().max()) <= 2e-2 * scale
    assert bool((dZ[:, E:] == 0).all())
    assert torch.allclose(dW.double().cpu(), ref_dw, rtol=2e-3, atol=2e-3 * float(ref_dw.abs().max() + 1e-6))
    assert torch.allclose(db.double().cpu(), ref_db, rtol=2e-3, atol=1e-4)
    assert torch.allclose(dbelow.double().cpu(), ref_dz.sum(0), rtol=2e-2, atol=2e-2 * scale)


def _programs_world(terminal, batch, n_max, ragged, seed, relate_prob=0.6, neg_rel=False):
    from dfol_vqa_b200 import synth
    from dfol_vqa_b200.ontology import synthetic_ontology
    from dfol_vqa_b200.programs import ProgramCollater
    dims = dict(box=2048, feat=512, hidden=256, emb=300)
    ont = synthetic_ontology(400, 60, 6, 5, seed=3, embedding_dim=300)
    if terminal == 'chain':
        questions = synth.make_relation_chain_questions(ont, batch, 5, seed=seed)
    else:
        questions = synth.make_questions(ont, batch, terminal, 1, 3, seed=seed, relate_prob=relate_prob)
    if neg_rel:   # negate the relation of every other relate hop (the others become round-trip predicates of their slot)
        k = 0
        for q in questions:
            for br in q['program']['branches']:
                for op in br:
                    if op['operator'] == 'relate':
                        if k % 2 == 0:
                            op['arguments'][0] = 'not(%s)' % op['arguments'][0]
                        k += 1
    counts = synth.object_counts(batch, n_max, ragged, seed=seed + 1)
    feats, bidx = synth.make_object_features(counts, 2048, seed=seed + 2)
    pbs = ProgramCollater(1, lambda qs: (feats, bidx)).collate(questions)
    return ont, dims, pbs


@pytest.mark.parametrize('ptab', [True, False])
@pytest.mark.parametrize('terminal,n_max,ragged,neg_rel', [
    ('chain', 48, False, False), ('verify_rel', 37, True, False), ('choose_rel', 24, True, False),
    ('query_attr', 48, False, False), ('and', 100, False, False),
    # option lists in probability space: long lists (several passes per warp), > 64 objects, ragged counts
    ('query_attr', 100, True, False), ('query_attr', 20, True, False), ('choose_attr', 37, True, False),
    # every hop geometry (8 / 16 / 32 lanes per tile row), aligned and ragged object counts, negated relations
    ('chain', 30, True, False), ('chain', 32, False, True), ('chain', 61, True, True), ('chain', 64, False, False),
    ('chain', 100, False, True), ('chain', 125, True, False), ('chain', 128, False, False), ('chain', 99, True, True)])
def test_fast_interpreter_matches_exact(terminal, n_max, ragged, neg_rel, ptab):
    """dfol_program_{fwd,bwd}_fast (bulk-async tile ring, probability-space relate hop with the transposed row
    reduction, MUFU math) against the exact kernels on the same tables: log-probabilities and the compact gradient
    slices.  ptab: with / without the probability table of the slot kernels."""
    from dfol_vqa_b200 import capi
    from dfol_vqa_b200.engine import SceneLayout
    ont, dims, pbs = _programs_world(terminal, 12, n_max, ragged, seed=41, neg_rel=neg_rel)
    # negated relations at the trained-like operating point (p ~ 0.02, not(p) ~ 0.98) put every chain into the regime
    # where 1 - prod(1 - m) cancels to a few fp32 ulps of 1 and BOTH kernels carry only ~2 digits; a moderate operating
    # point keeps those cases well conditioned so that the comparison means something
    interp = helpers.build_interpreter(ont, dims, seed=5, gemm_mode='bf16', emb_bias=-1.5 if neg_rel else -4.0)
    pb = pbs[0].to_cuda(0)
    cp = interp.compiled(pb, False)
    counts = interp._object_counts(pb)
    layout = SceneLayout.get(counts, interp._weights.emb.weight.shape[0], len(ont._relation_index), torch.device('cuda', 0))
    eng = interp._engine
    with torch.no_grad():
        scene = eng.build_scene(pb._object_features.float(), layout, keep_for_backward=True, cp=cp)
        if layout.P > 0:
            # the probability table is e^{ll} (zero on self pairs) wherever the slot kernels wrote a row
            assert scene.rel_p is not None
            ll, pp = scene.rel_ll, scene.rel_p
            for b, n in enumerate(counts):
                stride = int(layout.rel_stride[b])
                for j in range(int(cp.img_slot[b + 1] - cp.img_slot[b])):
                    off = int(cp.slot_blk[b]) + j * stride
                    l, q = ll[off:off + n * n].view(n, n), pp[off:off + n * n].view(n, n)
                    assert bool((q.diagonal() == 0).all()) and bool((l.diagonal() == -30.0).all())
                    mask = ~torch.eye(n, dtype=torch.bool, device=l.device)
                    assert torch.allclose(q[mask], l[mask].exp(), rtol=1e-5, atol=1e-30)
        if not ptab:
            scene.rel_p = None
        out = {}
        for mode in ('bf16', 'fp32'):   # 'bf16' -> *_fast entry points, 'fp32' -> exact entry points
            eng.gemm_mode = mode
            lp, tape = eng.run_programs(cp, scene, save_tape=True)
            d_lp = torch.linspace(-1.0, 1.0, lp.numel(), device='cuda')
            g_attr, g_rel = eng.program_backward(cp, scene, tape, d_lp)
            out[mode] = (lp.clone(), g_attr.clone(), g_rel.clone())
        eng.gemm_mode = 'bf16'
    lp_f, ga_f, gr_f = out['bf16']
    lp_e, ga_e, gr_e = out['fp32']
    assert bool(torch.isfinite(lp_f).all()) and bool(torch.isfinite(ga_f).all()) and bool(torch.isfinite(gr_f).all())
    assert float(lp_e.max()) > -40.0    # not a batch of clamped constants
    ok = (lp_f - lp_e).abs() <= 2e-3 * lp_e.abs() + 2e-4
    sat = (lp_f.exp() - lp_e.exp()).abs() <= 1e-6     # fp32 resolution of probabilities next to 0 / 1
    assert bool((ok | sat).all()), (lp_f, lp_e)
    for a, b in ((ga_f, ga_e), (gr_f, gr_e)):
        scale = float(b.abs().max())
        if scale < 1e3:   # saturated programs produce 1e12-scale BCE-like gradients: not comparable
            assert float((a - b).abs().max()) <= 2e-2 * scale + 1e-6, float((a - b).abs().max()) / (scale + 1e-9)


@pytest.mark.parametrize('terminal', ['verify_rel', 'choose_rel'])
def test_relation_slots_match_dense_table(terminal):
    """Demand-driven relation columns (compiler slots + slot kernels) give the same log-probabilities as the dense
    [nR] relation table of the same tensor-core scene build."""
    from dfol_vqa_b200.compiler import ProgramCompiler
    from dfol_vqa_b200.engine import SceneLayout
    ont, dims, pbs = _programs_world(terminal, 10, 20, True, seed=51)
    interp = helpers.build_interpreter(ont, dims, seed=5, gemm_mode='bf16', emb_bias=-4.0)
    pb = pbs[0].to_cuda(0)
    counts = interp._object_counts(pb)
    layout = SceneLayout.get(counts, interp._weights.emb.weight.shape[0], len(ont._relation_index), torch.device('cuda', 0))
    eng = interp._engine
    lps = []
    with torch.no_grad():
        for slots in (True, False):
            cp = ProgramCompiler(ont, normalize=True, relation_slots=slots).compile(pb, counts)
            for training in ((True, False) if slots else (True,)):
                scene = eng.build_scene(pb._object_features.float(), layout, keep_for_backward=training, cp=cp)
                lp, _ = eng.run_programs(cp, scene, save_tape=False)
                lps.append(lp.clone())
    for lp in lps[:-1]:
        assert torch.allclose(lp, lps[-1], rtol=2e-2, atol=2e-2), (lp - lps[-1]).abs().max()


def test_host_step_pipeline_matches_direct_steps():
    """HostStepPipeline (H2D staging on a copy stream, lagged read-back) returns the same per-step results, in order,
    as stepping the device batches directly."""
    from dfol_vqa_b200.pipeline import HostStepPipeline
    ont, dims, _ = _programs_world('exist', 8, 16, True, seed=61)
    interp = helpers.build_interpreter(ont, dims, seed=5, gemm_mode='bf16', emb_bias=-4.0)
    host = []
    for i in range(4):
        _, _, pbs = _programs_world('exist', 8, 16, True, seed=70 + i)
        host.append(pbs[0].pin_memory())

    def step(pb):
        with torch.no_grad():
            return interp([pb], True)['log_probability']

    direct = [step(hb.to_cuda(0)).cpu() for hb in host]
    piped = HostStepPipeline(step, torch.device('cuda', 0)).run(host)
    assert len(piped) == len(direct)
    for a, b in zip(piped, direct):
        assert torch.equal(a, b)


@pytest.mark.parametrize('workload,batch,n,mode', [('c1', 256, 48, 'bf16'), ('c1', 256, 48, 'fp32'), ('c3', 256, 100, 'bf16')])
def test_full_size_question_independence(workload, batch, n, mode):
    """Size-independent property at BASELINE.json's full sizes: every question only depends on its own image, so the
    log-probabilities of the first questions of the full batch are BIT-identical to running those questions alone
    (forward kernels have no cross-question reductions), and finite everywhere."""
    from dfol_vqa_b200 import synth
    from dfol_vqa_b200.ontology import synthetic_ontology
    from dfol_vqa_b200.programs import ProgramCollater
    dims = dict(box=2048, feat=512, hidden=256, emb=300)
    ont = synthetic_ontology(seed=1, embedding_dim=300, concept_num=2335, relation_num=333, category_num=31, class_num=53)
    interp = helpers.build_interpreter(ont, dims, seed=0, gemm_mode=mode, emb_bias=-4.0)
    if workload == 'c3':
        questions = synth.make_relation_chain_questions(ont, batch, 9, seed=5)
    else:
        questions = synth.make_questions(ont, batch, 'verify_rel', 1, 3, seed=5, relate_prob=0.35)
    feats, bidx = synth.make_object_features([n] * batch, 2048, seed=6)
    sub = 16
    full = ProgramCollater(1, lambda qs: (feats, bidx)).collate(questions)[0].to_cuda(0)
    part = ProgramCollater(1, lambda qs: (feats[:sub * n].clone(), bidx[:sub * n].clone())).collate(
        json.loads(json.dumps(questions[:sub])))[0].to_cuda(0)
    with torch.no_grad():
        lp_full = interp([full], True)['log_probability'].clone()
        lp_part = interp([part], True)['log_probability'].clone()
    assert lp_full.numel() == batch and bool(torch.isfinite(lp_full).all())
    assert bool((lp_full <= 1e-6).all())
    assert torch.equal(lp_full[:sub], lp_part), (lp_full[:sub] - lp_part).abs().max()


def test_full_size_gradient_additivity():
    """Size-independent property of the training step at the full c1 size: loss and gradients of the whole batch equal
    the accumulation over two sub-batches (the reference's split_num semantics, trainer.py:429-442)."""
    from dfol_vqa_b200 import synth
    from dfol_vqa_b200.interpreter import FusedTrainStep
    from dfol_vqa_b200.ontology import synthetic_ontology
    from dfol_vqa_b200.programs import ProgramCollater
    dims = dict(box=2048, feat=512, hidden=256, emb=300)
    batch, n = 256, 48
    ont = synthetic_ontology(seed=1, embedding_dim=300, concept_num=2335, relation_num=333, category_num=31, class_num=53)
    interp = helpers.build_interpreter(ont, dims, seed=0, gemm_mode='bf16', emb_bias=-4.0)
    questions = synth.make_questions(ont, batch, 'verify_rel', 1, 3, seed=15, relate_prob=0.35)
    feats, bidx = synth.make_object_features([n] * batch, 2048, seed=16)
    full = ProgramCollater(1, lambda qs: (feats, bidx)).collate(json.loads(json.dumps(questions)))
    halves = ProgramCollater(2, helpers.slicing_source(feats, bidx)).collate(json.loads(json.dumps(questions)))
    assert len(halves) == 2
    step = FusedTrainStep(interp)
    loss_full = float(step.forward_backward(helpers.to_cuda(full), global_question_num=batch))
    g_full = step.flat_grad.clone()
    loss_split = float(step.forward_backward(helpers.to_cuda(halves), global_question_num=batch))
    g_split = step.flat_grad.clone()
    assert abs(loss_full - loss_split) <= 1e-4 * max(1.0, abs(loss_full))
    scale = float(g_full.abs().max())
    assert scale > 0 and bool(torch.isfinite(g_full).all())
    assert float((g_full - g_split).abs().max()) <= 2e-3 * scale


def test_bf16_staged_features_are_bit_identical():
    """ProgramBatch.stage_bf16 (bf16 box features + fp32 geometry on the host, half the H2D bytes): the tensor-core scene
    build casts the fp32 features to bf16 as its first step, so log-probabilities and loss are BIT-identical (gradients up to the order of their atomic reductions)."""
    import copy
    from dfol_vqa_b200.interpreter import FusedTrainStep
    ont, dims, pbs = _programs_world('verify_rel', 12, 24, True, seed=91)
    interp = helpers.build_interpreter(ont, dims, seed=5, gemm_mode='bf16', emb_bias=-4.0)
    plain = pbs[0]
    staged = copy.copy(plain).stage_bf16(drop_fp32=True)
    assert staged._object_features is None and plain._object_features is not None
    h2d_plain = plain._object_features.numel() * 4
    h2d_staged = sum(t.numel() * t.element_size() for t in staged._staged)
    assert h2d_staged < 0.51 * h2d_plain
    step = FusedTrainStep(interp)
    interp.train()
    out = []
    for pb in (plain, staged):
        dev_pb = pb.pin_memory().to_cuda(0)
        with torch.no_grad():
            lp = interp([dev_pb], True)['log_probability'].clone()
        loss = step.forward_backward([dev_pb]).clone()
        out.append((lp, loss, step.flat_grad.clone()))
    assert torch.equal(out[0][0], out[1][0])
    assert torch.equal(out[0][1], out[1][1])
    # (the weight gradients are split-K reductions with atomic adds: equal up to summation order)
    scale = float(out[0][2].abs().max())
    assert float((out[0][2] - out[1][2]).abs().max()) <= 1e-5 * scale
    fp32 = helpers.build_interpreter(ont, dims, seed=5, gemm_mode='fp32', emb_bias=-4.0)
    with pytest.raises(RuntimeError):
        fp32([staged.to_cuda(0)], True)


@pytest.mark.parametrize('terminal,n_max,ragged', [('chain', 48, False), ('chain', 37, True), ('verify_rel', 100, False),
                                                   ('choose_rel', 5, True)])
def test_rel_slots_tensor_core_kernel_matches_simt(terminal, n_max, ragged, monkeypatch):
    """dfol_rel_slots_fwd_tc (grouped tcgen05 GEMM: every image against its own slot rows, bf16 weights) against
    dfol_rel_slots_fwd (fp32 weights, SIMT dot products) on the same stored activation: the compact relation tables."""
    from dfol_vqa_b200.engine import SceneLayout
    ont, dims, pbs = _programs_world(terminal, 12, n_max, ragged, seed=43)
    interp = helpers.build_interpreter(ont, dims, seed=5, gemm_mode='bf16', emb_bias=-4.0)
    pb = pbs[0].to_cuda(0)
    cp = interp.compiled(pb, False)
    counts = interp._object_counts(pb)
    layout = SceneLayout.get(counts, interp._weights.emb.weight.shape[0], len(ont._relation_index), torch.device('cuda', 0))
    tables = {}
    for name, thr in (('tc', '1'), ('simt', '99')):
        monkeypatch.setenv('DFOL_SLOTS_TC_MIN', thr)
        with torch.no_grad():
            scene = interp._engine.build_scene(pb._object_features.float(), layout, keep_for_backward=True, cp=cp)
        torch.cuda.synchronize()
        tables[name] = scene.rel_ll.clone()
    # valid entries: n_b^2 per slot (slices are padded to a multiple of 4 floats; the padding is never read)
    valid = torch.zeros(tables['tc'].numel(), dtype=torch.bool)
    for b, n in enumerate(counts):
        stride = int(layout.rel_stride[b])
        for j in range(int(cp.img_slot[b + 1] - cp.img_slot[b])):
            off = int(cp.slot_blk[b]) + j * stride
            valid[off:off + n * n] = True
    valid = valid.cuda()
    a, b = tables['tc'][valid], tables['simt'][valid]
    assert a.numel() > 0 and bool(torch.isfinite(a).all())
    off_diag = b != -30.0
    assert bool(((a == -30.0) == (b == -30.0)).all())   # self pairs agree exactly
    assert float((a - b)[off_diag].abs().max()) <= 2e-2 + 2e-2 * float(b[off_diag].abs().max())


@pytest.mark.parametrize('counts', [[48] * 6, [5, 9, 3, 7, 1, 2], [100, 37, 125], [64, 63, 65]])
def test_pair_hidden_mma_matches_simt(counts):
    """dfol_pair_hidden_fwd_mma (one-hot grouped tcgen05 GEMM: U, V, Wg and the bias as bf16 operands) against
    dfol_pair_hidden_fwd_tc (fp32 sums, one bf16 rounding) on the same U|V: the bf16 hidden layer and the geometry table."""
    from dfol_vqa_b200.capi import call, ptr, stream_ptr
    from dfol_vqa_b200.engine import SceneLayout
    H, F = 256, 32
    T, P = sum(counts), sum(c * c for c in counts)
    g = torch.Generator().manual_seed(sum(counts))
    uv = (torch.randn(T, 2 * H, generator=g) * 0.7).cuda()
    obj = torch.rand(T, F + 4, generator=g).cuda()
    wg = (torch.randn(H, 4, generator=g) * 0.5).cuda()
    bias = (torch.randn(H, generator=g) * 0.3).cuda()
    layout = SceneLayout.get(counts, 10, 4, torch.device('cuda', 0))
    outs = {}
    for name in ('simt', 'mma'):
        h = torch.full((P, H), float('nan'), device='cuda', dtype=torch.bfloat16)
        geo = torch.full((P, 4), float('nan'), device='cuda')
        args = (ptr(uv), 2 * H, ptr(obj[:, F:]), F + 4, ptr(wg), 4, ptr(bias), ptr(h), H, H, ptr(geo),
                ptr(layout.pair_row), ptr(layout.obj_row), ptr(layout.img_n), len(counts), max(counts))
        if name == 'mma':
            ws = torch.empty(len(counts) * H, (2 * max(counts) + 5 + 63) // 64 * 64, device='cuda', dtype=torch.bfloat16)
            call('dfol_pair_hidden_fwd_mma', *args, ptr(ws), stream_ptr())
        else:
            call('dfol_pair_hidden_fwd_tc', *args, stream_ptr())
        torch.cuda.synchronize()
        outs[name] = (h.float(), geo)
    (ha, ga), (hb, gb) = outs['mma'], outs['simt']
    assert bool(torch.isfinite(ha).all())
    assert torch.equal(ga, gb)
    # operands rounded to bf16 before the sum (|U|, |V| ~ 2): a few bf16 ulps of the operands
    assert float((ha - hb).abs().max()) <= 4e-2, float((ha - hb).abs().max())
    assert float((ha - hb).abs().mean()) <= 4e-3
