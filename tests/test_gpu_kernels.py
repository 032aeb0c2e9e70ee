"""Kernel-level GPU tests through the C ABI: GEMMs (fp32 SIMT and bf16 tcgen05), loss, Adam, bf16-mode forward."""

import json

import numpy as np
import pytest
import torch

import helpers
import dfol_oracle as orc

pytestmark = pytest.mark.gpu


def _table_maps(counts, ncols, pairs, dev):
    n = np.asarray(counts, dtype=np.int64)
    rows = n * n if pairs else n
    stride = (rows + 3) // 4 * 4
    row0 = np.concatenate([[0], np.cumsum(rows)])
    blk = ncols * np.concatenate([[0], np.cumsum(stride)])
    t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a.astype(d))).to(dev)
    maps = {'row_img': torch.repeat_interleave(torch.arange(len(counts), dtype=torch.int32), torch.from_numpy(rows)).to(dev),
            'img_row': t(row0, np.int32), 'img_blk': t(blk[:-1], np.int64), 'img_stride': t(stride, np.int32)}
    if pairs:
        maps['img_n'] = t(n, np.int32)
        maps['diag'] = -30.0
    return maps, int(row0[-1]), int(blk[-1]), row0, blk, stride


def _act(x, act):
    if act == 1:
        return torch.nn.functional.elu(x)
    if act == 2:
        return torch.sigmoid(x)
    if act == 3:
        return torch.nn.functional.logsigmoid(x)
    return x


@pytest.mark.parametrize('M,N,K', [(1, 1, 1), (130, 70, 33), (257, 300, 256), (1000, 333, 300), (64, 512, 2048)])
@pytest.mark.parametrize('act', [0, 1, 2, 3])
def test_gemm_f32_forward(M, N, K, act):
    from dfol_vqa_b200.engine import gemm_f32
    g = torch.Generator().manual_seed(M * 7 + N)
    A = torch.randn(M, K + 3, generator=g)[:, :K]
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    C = torch.empty(M, N, device='cuda')
    gemm_f32(A.cuda(), W.cuda().t(), C, b.cuda(), act)
    ref = _act(A.double() @ W.double().t() + b.double(), act)
    tol = 2e-5 * max(1.0, (K / 256) ** 0.5)
    assert torch.allclose(C.cpu().double(), ref, rtol=tol, atol=tol)


def test_gemm_f32_variants():
    from dfol_vqa_b200.engine import gemm_f32
    from dfol_vqa_b200.capi import K as KK
    g = torch.Generator().manual_seed(3)
    M, N, Kd = 300, 130, 5000
    dZ = torch.randn(Kd, M, generator=g).cuda()      # wgrad: C = dZ^T @ X, K = rows
    X = torch.randn(Kd, N, generator=g).cuda()
    C = torch.zeros(M, N, device='cuda')
    gemm_f32(dZ.t(), X, C, split_k=8)
    ref = dZ.double().t() @ X.double()
    assert torch.allclose(C.double(), ref, rtol=1e-4, atol=1e-3)
    C2 = torch.ones(M, N, device='cuda')
    gemm_f32(dZ.t(), X, C2, accumulate=True)
    assert torch.allclose(C2.double(), ref + 1.0, rtol=1e-4, atol=1e-3)
    # dgrad with fused activation-derivative multiplier
    W = torch.randn(N, 77, generator=g).cuda()
    Hs = torch.rand(Kd, 77, generator=g).cuda() - 0.3
    D = torch.empty(Kd, 77, device='cuda')
    gemm_f32(X, W, D, mul_src=Hs, mul_mode=KK.MUL_ELU_GRAD)
    ref = (X.double() @ W.double()) * torch.where(Hs > 0, torch.ones_like(Hs), Hs + 1).double()
    assert torch.allclose(D.double(), ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('pairs', [False, True])
@pytest.mark.parametrize('impl', ['f32', 'tc'])
def test_gemm_table_store(pairs, impl):
    """Per-image transposed table store (+ -30 on self pairs) of both GEMM kernels."""
    from dfol_vqa_b200.engine import gemm_f32
    from dfol_vqa_b200.engine_tc import TensorCorePath as ReasoningEngine
    counts = [5, 12, 3, 9, 16]
    ncols, Kd = 37, 128
    dev = torch.device('cuda')
    maps, M, size, row0, blk, stride = _table_maps(counts, ncols, pairs, dev)
    g = torch.Generator().manual_seed(11)
    A = torch.randn(M, Kd, generator=g)
    W = torch.randn(ncols, Kd, generator=g) / Kd ** 0.5
    b = torch.randn(ncols, generator=g)
    out = torch.full((size,), float('nan'), device=dev)
    if impl == 'f32':
        gemm_f32(A.cuda(), W.cuda().t(), out, b.cuda(), 3, table=maps)
        ref = _act(A.double() @ W.double().t() + b.double(), 3)
        tol = dict(rtol=2e-5, atol=2e-6)
    else:
        A16, W16 = A.cuda().bfloat16(), W.cuda().bfloat16()
        ReasoningEngine._tc(A16, W16, out, ncols, Kd, b.cuda(), 3, None, table=maps)
        ref = _act(A16.cpu().double() @ W16.cpu().double().t() + b.double(), 3)
        tol = dict(rtol=2e-3, atol=2e-3)
    out = out.cpu().double()
    for i, n in enumerate(counts):
        rows = n * n if pairs else n
        block = out[int(blk[i]):int(blk[i]) + ncols * int(stride[i])].view(ncols, int(stride[i]))[:, :rows].t()
        want = ref[int(row0[i]):int(row0[i]) + rows].clone()
        if pairs:
            diag = torch.arange(n) * n + torch.arange(n)
            want[diag] = -30.0
        assert torch.allclose(block, want, **tol), (i, (block - want).abs().max())


@pytest.mark.parametrize('M,N,K', [(128, 16, 64), (300, 300, 256), (1000, 333, 320), (77, 512, 576), (4608, 2335, 320)])
@pytest.mark.parametrize('act,out16', [(0, False), (2, True), (1, True), (3, False)])
def test_gemm_bf16_tcgen05(M, N, K, act, out16):
    """tcgen05/TMA GEMM vs an fp64 product of the same bf16 operands; padding columns of C must be zero."""
    from dfol_vqa_b200.engine_tc import TensorCorePath as ReasoningEngine
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g)).cuda().bfloat16()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda().bfloat16()
    b = torch.randn(N, generator=g).cuda()
    ldc = (N + 63) // 64 * 64
    C = torch.full((M, ldc), float('nan'), device='cuda', dtype=torch.bfloat16 if out16 else torch.float32)
    ReasoningEngine._tc(A, W, C, N, K, b, act, None)
    torch.cuda.synchronize()
    ref = _act(A.cpu().double() @ W.cpu().double().t() + b.cpu().double(), act)
    got = C.cpu().double()
    tol = dict(rtol=1.5e-2, atol=1.5e-2) if out16 else dict(rtol=2e-3, atol=2e-3)
    assert torch.allclose(got[:, :N], ref, **tol), (got[:, :N] - ref).abs().max()
    assert bool((got[:, N:] == 0).all())


@pytest.mark.parametrize('kind', [0, 1, 2])
def test_loss_kernel(kind):
    from dfol_vqa_b200.capi import call, ptr, stream_ptr
    g = torch.Generator().manual_seed(5 + kind)
    n = 37
    lp = (-torch.rand(n, generator=g) * 4).requires_grad_(True)
    seg = [0, 3, 5, 12, 20, 37]
    if kind == 0:
        target = (torch.rand(n, generator=g) > 0.5).float()
        ref = torch.nn.functional.binary_cross_entropy(lp.exp(), target, reduction='sum')
    elif kind == 1:
        target = torch.zeros(n)
        target[[1, 4, 7, 15, 30]] = 1
        ref = sum(orc.safe_log(lp[a:b].exp().sum()) for a, b in zip(seg[:-1], seg[1:])) - (target * lp).sum()
    else:
        target = torch.zeros(n)
        ref = -lp.sum()
    scale = 1.0 / 6
    (ref * scale).backward()
    out = torch.zeros(1, device='cuda')
    dlp = torch.empty(n, device='cuda')
    segd = torch.tensor(seg, dtype=torch.int32, device='cuda')
    lpd, td = lp.detach().cuda(), target.cuda()   # keep the operands alive until the stream has consumed them
    call('dfol_loss_fwd_bwd', ptr(lpd), ptr(td), ptr(segd), len(seg) - 1, n, kind, scale, ptr(out), ptr(dlp),
         stream_ptr())
    assert abs(float(out) - float(ref) * scale) <= 1e-5 * abs(float(ref) * scale) + 1e-7
    assert torch.allclose(dlp.cpu(), lp.grad, rtol=1e-5, atol=1e-7)


def test_clip_adam_matches_torch():
    """clip_grad_norm_(0.65) + torch.optim.Adam(lr, weight_decay) for three steps (reference trainer.py:438-441)."""
    from dfol_vqa_b200.capi import call, ptr, stream_ptr
    g = torch.Generator().manual_seed(9)
    p0 = torch.randn(5000, generator=g)
    p_ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=1e-2, weight_decay=1e-3)
    p = p0.clone().cuda()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        grad = torch.randn(5000, generator=g) * (3.0 if step == 2 else 0.001)
        p_ref.grad = grad.clone()
        torch.nn.utils.clip_grad_norm_([p_ref], 0.65)
        opt.step()
        ss = torch.zeros(1, device='cuda')
        gd = grad.cuda()
        call('dfol_sumsq', ptr(gd), gd.numel(), ptr(ss), stream_ptr())
        call('dfol_adam_step', ptr(p), ptr(gd), ptr(m), ptr(v), p.numel(), ptr(ss), 0.65, 1e-2, 0.9, 0.999, 1e-8, 1e-3,
             step, stream_ptr())
        assert torch.allclose(p.cpu(), p_ref.detach(), rtol=1e-5, atol=1e-6), step


@pytest.mark.parametrize('terminal', ['exist', 'verify_rel', 'query_attr', 'choose_rel'])
def test_bf16_mode_answer_logits(terminal):
    """bf16-GEMM mode (tcgen05 scene build): answer logits within 2e-2 of the fp32 oracle and identical argmax
    answers (north_star tolerance)."""
    from dfol_vqa_b200 import synth
    from dfol_vqa_b200.ontology import synthetic_ontology
    from dfol_vqa_b200.programs import ProgramCollater
    dims = dict(box=2048, feat=512, hidden=256, emb=300)
    ont = synthetic_ontology(400, 60, 6, 5, seed=3, embedding_dim=300)
    interp = helpers.build_interpreter(ont, dims, seed=5, gemm_mode='bf16')
    questions = synth.make_questions(ont, 12, terminal, 1, 4, seed=21)
    counts = synth.object_counts(12, 40, True, seed=22)
    feats, bidx = synth.make_object_features(counts, 2048, seed=23)
    pbs = ProgramCollater(1, lambda qs: (feats, bidx)).collate(questions)
    params = helpers.oracle_params(interp)
    with torch.no_grad():
        results, _ = orc.run_step(ont, params, ProgramCollater(1, lambda qs: (feats, bidx)).collate(
            json.loads(json.dumps(questions))), is_training=False)
    interp.eval()
    with torch.no_grad():
        out = interp(helpers.to_cuda(pbs), False)
    lp = out['log_probability'].cpu()
    ref = results[0]['log_probability']
    assert (lp - ref).abs().max() <= 2e-2 * max(1.0, float(ref.abs().max())), float((lp - ref).abs().max())
    # identical answers wherever the oracle's decision margin exceeds the bf16 tolerance
    p = ref.exp()
    checked = 0
    if out['type'] == 0:
        for q, (a, b) in enumerate(zip(out['answer'], results[0]['answer'])):
            if abs(float(p[q]) - 0.5) > 0.05:
                assert a == b, (q, a, b)
                checked += 1
    else:
        start = 0
        for q, opts in enumerate(results[0]['options']):
            seg = p[start:start + len(opts)].sort(descending=True)[0]
            start += len(opts)
            if len(opts) == 1 or float(seg[0] - seg[1]) > 0.05 * float(seg[0]):
                assert sorted(out['answer'][q]) == sorted(results[0]['answer'][q]), q
                checked += 1
    assert checked > 0


@pytest.mark.parametrize('M,N,K', [(128, 64, 64), (300, 256, 5000), (512, 2048, 12288), (100, 516, 777), (300, 256, 589824)])
def test_gemm_bf16_tcgen05_wgrad(M, N, K):
    """MN-major split-K tcgen05 contraction C += A^T B (reduction over rows) vs fp64 on the same bf16 operands."""
    from dfol_vqa_b200.capi import call, ptr, stream_ptr
    g = torch.Generator().manual_seed(M + N)
    lda, ldb = (M + 63) // 64 * 64, (N + 7) // 8 * 8
    A = torch.zeros(K, lda, dtype=torch.bfloat16)
    A[:, :M] = (torch.randn(K, M, generator=g) * 0.5).bfloat16()
    B = torch.zeros(K, ldb, dtype=torch.bfloat16)
    B[:, :N] = (torch.randn(K, N, generator=g) * 0.5).bfloat16()
    Ad, Bd = A.cuda(), B.cuda()
    C = torch.ones(M, N, device='cuda')
    call('dfol_gemm_bf16_tc_wgrad', ptr(Ad), lda, ptr(Bd), ldb, ptr(C), N, M, N, K, stream_ptr())
    torch.cuda.synchronize()
    ref = Ad[:, :M].double().t() @ Bd[:, :N].double() + 1.0
    err = (C.double() - ref).abs().max()
    assert err <= 1e-3 * (K ** 0.5) * 0.25 + 1e-3, float(err)


@pytest.mark.parametrize('M,N,K', [(1000, 256, 320), (4608, 256, 320), (300, 64, 64)])
@pytest.mark.parametrize('mode', [0, 1, 2])
def test_gemm_bf16_tcgen05_dgrad(M, N, K, mode):
    from dfol_vqa_b200.capi import call, ptr, stream_ptr
    g = torch.Generator().manual_seed(M + mode)
    dZ = (torch.randn(M, K, generator=g)).cuda().bfloat16()
    Wt = (torch.randn(N, K, generator=g) / K ** 0.5).cuda().bfloat16()
    Hs = (torch.rand(M, N, generator=g) - 0.3).cuda().bfloat16()
    dX = torch.full((M, N), float('nan'), device='cuda', dtype=torch.bfloat16)
    call('dfol_gemm_bf16_tc_dgrad', ptr(dZ), K, ptr(Wt), K, ptr(dX), N, 0, M, N, K, ptr(Hs), N, mode, 1.0, stream_ptr())
    torch.cuda.synchronize()
    ref = dZ.double() @ Wt.double().t()
    h = Hs.double()
    if mode == 1:
        ref = ref * h * (1 - h)
    elif mode == 2:
        ref = ref * torch.where(h > 0, torch.ones_like(h), h + 1)
    assert torch.allclose(dX.double(), ref, rtol=1.5e-2, atol=1.5e-2), (dX.double() - ref).abs().max()


@pytest.mark.parametrize('terminal', ['verify_rel', 'exist', 'and', 'query_attr'])
def test_bf16_mode_training_gradients(terminal):
    """Training step in bf16 tensor-core mode: loss within 2e-2, gradients within a few percent of each tensor's
    scale of the fp32 oracle (mixed precision: bf16 operands, fp32 accumulation / master weights)."""
    from dfol_vqa_b200 import synth
    from dfol_vqa_b200.ontology import synthetic_ontology
    from dfol_vqa_b200.programs import ProgramCollater
    from dfol_vqa_b200.interpreter import FusedTrainStep
    dims = dict(box=2048, feat=512, hidden=256, emb=300)
    ont = synthetic_ontology(400, 60, 6, 5, seed=3, embedding_dim=300)
    interp = helpers.build_interpreter(ont, dims, seed=5, gemm_mode='bf16', emb_bias=-4.0)
    questions = synth.make_questions(ont, 16, terminal, 1, 3, seed=31, relate_prob=0.6)
    counts = synth.object_counts(16, 48, True, seed=32)
    feats, bidx = synth.make_object_features(counts, 2048, seed=33)
    pbs = ProgramCollater(1, lambda qs: (feats, bidx)).collate(questions)
    params = helpers.oracle_params(interp, torch.float32, requires_grad=True)
    results, loss_ref = orc.run_step(ont, params, ProgramCollater(1, lambda qs: (feats, bidx)).collate(
        json.loads(json.dumps(questions))), is_training=True)
    loss_ref.backward()
    step = FusedTrainStep(interp)
    loss = step.forward_backward(helpers.to_cuda(pbs))
    assert abs(float(loss) - float(loss_ref)) <= 2e-2 * max(1.0, abs(float(loss_ref)))
    keys = {id(p): k for k, p in interp.named_parameters()}
    for p in interp.oracle_parameters():
        k = keys[id(p)]
        g_ref = params[k].grad if params[k].grad is not None else torch.zeros_like(params[k])
        scale = float(g_ref.abs().max())
        err = float((step.grads[id(p)].cpu() - g_ref).abs().max())
        assert err <= 6e-2 * scale + 1e-6, (k, err, scale)


@pytest.mark.parametrize('M,N,K,act', [(1000, 300, 256, 2), (4608 + 77, 300, 256, 2), (300, 256, 320, 1), (129, 64, 64, 0),
                                       (40000, 300, 256, 2)])
@pytest.mark.parametrize('impl', ['resident', 'cluster'])
def test_pair_layer_fwd_resident_gemm(M, N, K, act, impl):
    """Persistent weights-resident tcgen05 GEMMs (single CTA and cluster-of-two): bf16 output with zero K-padding
    vs fp64."""
    from dfol_vqa_b200.capi import call, ptr, stream_ptr
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 0.5).cuda().bfloat16()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda().bfloat16()
    b = torch.randn(N, generator=g).cuda()
    ldc = (N + 63) // 64 * 64
    C = torch.full((M, ldc), float('nan'), device='cuda', dtype=torch.bfloat16)
    if impl == 'cluster':
        call('dfol_pair_layer_fwd_cluster', ptr(A), K, ptr(W), K, ptr(C), ldc, ldc, ptr(b), M, N, K, act, stream_ptr())
    else:
        call('dfol_pair_layer_fwd_tc', ptr(A), K, ptr(W), K, ptr(C), ldc, ldc, ptr(b), M, N, K, act, None, 0, None,
             None, None, 0, None, None, None, None, None, 0.0, None, stream_ptr())
    torch.cuda.synchronize()
    z = A.double() @ W.double().t() + b.double()
    ref = {0: z, 1: torch.nn.functional.elu(z), 2: torch.sigmoid(z)}[act]
    assert torch.allclose(C[:, :N].double(), ref, rtol=1.5e-2, atol=1.5e-2), (C[:, :N].double() - ref).abs().max()
    assert bool((C[:, N:] == 0).all())


@pytest.mark.parametrize('counts,slots_per_image', [([5, 12, 3, 9, 16], [1, 0, 2, 4, 3]), ([48] * 6, [2, 1, 9, 0, 5, 4])])
@pytest.mark.parametrize('store', [True, False])
def test_pair_layer_fwd_slot_epilogue(counts, slots_per_image, store):
    """Layer-2 GEMM + demand-driven relation columns in the epilogue vs an fp64 evaluation from the same bf16
    operands; also checks the stand-alone dfol_rel_slots_fwd kernel on the stored activation."""
    from dfol_vqa_b200.capi import call, ptr, stream_ptr
    g = torch.Generator().manual_seed(sum(counts))
    K, E, C = 256, 300, 500
    n = torch.tensor(counts)
    P = int((n * n).sum())
    A = (torch.randn(P, K, generator=g) * 0.5).cuda().bfloat16()
    W2 = (torch.randn(E, K, generator=g) / K ** 0.5).cuda().bfloat16()
    b2 = torch.randn(E, generator=g).cuda()
    We = (torch.randn(C, E, generator=g) * 0.3).cuda()
    be = torch.randn(C, generator=g).cuda()
    stride = (n * n + 3) // 4 * 4
    img_slot = torch.tensor([0] + list(torch.tensor(slots_per_image).cumsum(0)), dtype=torch.int32)
    total_slots = int(img_slot[-1])
    slot_wrow = torch.randint(0, C, (max(total_slots, 1),), generator=g, dtype=torch.int32)
    blk = torch.tensor([0] + list((torch.tensor(slots_per_image) * stride).cumsum(0)), dtype=torch.int64)
    size = int(blk[-1])
    pair_row = torch.tensor([0] + list((n * n).cumsum(0)), dtype=torch.int32)
    row_img = torch.repeat_interleave(torch.arange(len(counts), dtype=torch.int32), n * n)
    dev = lambda t: t.cuda()
    d = dict(slot_wrow=dev(slot_wrow), img_slot=dev(img_slot), blk=dev(blk[:-1].contiguous()),
             stride=dev(stride.int()), row_img=dev(row_img), pair_row=dev(pair_row), img_n=dev(n.int()),
             img_nn=dev((n * n).int()))
    ldc = 320
    H2 = torch.full((P, ldc), float('nan'), device='cuda', dtype=torch.bfloat16) if store else None
    ll = torch.full((max(size, 1),), float('nan'), device='cuda')
    call('dfol_pair_layer_fwd_tc', ptr(A), K, ptr(W2), K, ptr(H2), ldc, ldc, ptr(b2), P, E, K, 2, ptr(We), E, ptr(be),
         ptr(d['slot_wrow']), ptr(d['img_slot']), max(slots_per_image), ptr(d['blk']), ptr(d['stride']),
         ptr(d['row_img']), ptr(d['pair_row']), ptr(d['img_n']), -30.0, ptr(ll), stream_ptr())
    torch.cuda.synchronize()
    h2 = torch.sigmoid(A.double() @ W2.double().t() + b2.double())
    ref = torch.full((max(size, 1),), float('nan'), dtype=torch.float64)
    h2c = h2.cpu()
    for b_i, nb in enumerate(counts):
        rows = h2c[int(pair_row[b_i]):int(pair_row[b_i + 1])]
        for k in range(slots_per_image[b_i]):
            wr = int(slot_wrow[int(img_slot[b_i]) + k])
            z = torch.nn.functional.logsigmoid(rows @ We[wr].double().cpu() + be[wr].double().cpu())
            z = z.view(nb, nb).clone()
            z.fill_diagonal_(-30.0)
            o = int(blk[b_i]) + k * int(stride[b_i])
            ref[o:o + nb * nb] = z.reshape(-1)
    mask = ~torch.isnan(ref)
    got = ll.double().cpu()
    assert torch.allclose(got[mask], ref[mask], rtol=2e-2, atol=2e-2), (got[mask] - ref[mask]).abs().max()
    if store:
        assert torch.allclose(H2[:, :E].double(), h2, rtol=1.5e-2, atol=1.5e-2)
        assert bool((H2[:, E:] == 0).all())
        ll2 = torch.full_like(ll, float('nan'))
        pp2 = torch.full_like(ll, float('nan'))
        call('dfol_rel_slots_fwd', ptr(H2), ldc, E, ptr(We), E, ptr(be), ptr(d['slot_wrow']), ptr(d['img_slot']),
             max(slots_per_image), ptr(d['blk']), ptr(d['stride']), ptr(d['pair_row']), ptr(d['img_nn']),
             ptr(d['img_n']), len(counts), max(counts) ** 2, -30.0, ptr(ll2), ptr(pp2), stream_ptr())
        torch.cuda.synchronize()
        got2 = ll2.double().cpu()
        assert torch.allclose(got2[mask], ref[mask], rtol=2e-2, atol=2e-2), (got2[mask] - ref[mask]).abs().max()
        # the probability table: e^{ll}, zero on self pairs (ll = -30)
        p2 = pp2.double().cpu()
        diag = mask & (got2 == -30.0)
        assert bool((p2[diag] == 0).all())
        off = mask & ~diag
        assert torch.allclose(p2[off], got2[off].exp(), rtol=1e-5, atol=1e-30)


@pytest.mark.parametrize('M,N,K', [(1000, 256, 320), (4608 + 5, 256, 320), (300, 64, 64), (40000, 256, 320)])
@pytest.mark.parametrize('mode', [0, 2])
@pytest.mark.parametrize('impl', ['resident', 'cluster'])
def test_pair_layer_dgrad_resident_gemm(M, N, K, mode, impl):
    from dfol_vqa_b200.capi import call, ptr, stream_ptr
    g = torch.Generator().manual_seed(M + mode)
    dZ = (torch.randn(M, K, generator=g)).cuda().bfloat16()
    Wt = (torch.randn(N, K, generator=g) / K ** 0.5).cuda().bfloat16()
    Hs = (torch.rand(M, N, generator=g) - 0.3).cuda().bfloat16()
    dX = torch.full((M, N), float('nan'), device='cuda', dtype=torch.bfloat16)
    entry = 'dfol_pair_layer_dgrad_cluster' if impl == 'cluster' else 'dfol_pair_layer_dgrad_tc'
    keep = (1.0,) if impl == 'cluster' else ()   # (the cluster entry point takes the dropout keep factor)
    call(entry, ptr(dZ), K, ptr(Wt), K, ptr(dX), N, 0, M, N, K, ptr(Hs), N, mode, *keep, stream_ptr())
    torch.cuda.synchronize()
    ref = dZ.double() @ Wt.double().t()
    h = Hs.double()
    if mode == 2:
        ref = ref * torch.where(h > 0, torch.ones_like(h), h + 1)
    assert torch.allclose(dX.double(), ref, rtol=1.5e-2, atol=1.5e-2), (dX.double() - ref).abs().max()


@pytest.mark.parametrize('M,N,K,k_real', [(100, 256, 320, 300), (129, 256, 320, 300), (1000, 256, 320, 300),
                                          (4608 * 3 + 5, 256, 320, 300), (40000, 192, 256, 256),
                                          (300000, 256, 320, 300)])
@pytest.mark.parametrize('mode,keep', [(2, 1.0), (1, 1.0), (2, 0.9)])
def test_pair_layer_dgrad_with_fused_weight_gradient(M, N, K, k_real, mode, keep):
    """dfol_pair_layer_dgrad_wgrad_cluster: the dgrad result is bit-identical to the plain cluster dgrad and the fused
    weight gradient equals dZ^T . h_saved (fp64 statement of nn.Linear's backward) and the stand-alone wgrad kernel."""
    from dfol_vqa_b200.capi import call, ptr, stream_ptr
    g = torch.Generator().manual_seed(M + mode)
    dZ = torch.randn(M, K, generator=g)
    dZ[:, k_real:] = 0
    dZ = dZ.cuda().bfloat16()
    Wt = (torch.randn(N, K, generator=g) / K ** 0.5).cuda().bfloat16()
    Hs = torch.rand(M, N, generator=g) - (0.3 if mode == 2 else 0.0)
    if keep < 1.0:
        Hs = Hs * (torch.rand(M, N, generator=g) < keep) / keep
    Hs = Hs.cuda().bfloat16()
    dX0 = torch.full((M, N), float('nan'), device='cuda', dtype=torch.bfloat16)
    dX1 = torch.full((M, N), float('nan'), device='cuda', dtype=torch.bfloat16)
    base = torch.randn(k_real, N, generator=g).cuda()
    dW, dW_sep = base.clone(), base.clone()
    call('dfol_pair_layer_dgrad_cluster', ptr(dZ), K, ptr(Wt), K, ptr(dX0), N, 0, M, N, K, ptr(Hs), N, mode, keep,
         stream_ptr())
    call('dfol_pair_layer_dgrad_wgrad_cluster', ptr(dZ), K, ptr(Wt), K, ptr(dX1), N, 0, M, N, K, ptr(Hs), N, mode, keep,
         ptr(dW), N, k_real, stream_ptr())
    call('dfol_gemm_bf16_tc_wgrad', ptr(dZ), K, ptr(Hs), N, ptr(dW_sep), N, k_real, N, M, stream_ptr())
    torch.cuda.synchronize()
    assert torch.equal(dX0, dX1)
    ref = base.double().cpu() + dZ.double().cpu()[:, :k_real].t() @ Hs.double().cpu()
    scale = float((ref - base.double().cpu()).abs().max())
    assert float((dW.double().cpu() - ref).abs().max()) <= 1e-4 * scale + 1e-4
    assert float((dW - dW_sep).abs().max()) <= 1e-4 * scale + 1e-4
