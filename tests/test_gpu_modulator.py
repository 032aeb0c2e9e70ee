"""Attention-transfer calibrator on the GPU: the interpreter kernels apply the modulations in
every hop (forward) and return d loss / d modulations (backward); held to fixtures recorded from the unmodified
reference with ``activate_attention_transfer: True`` and to the CPU oracle.  Needs a GPU."""

import pytest
import torch

import helpers
import dfol_oracle as orc

pytestmark = pytest.mark.gpu

FILES = helpers.golden_mod_files()


def _ids(p):
    return p.split('goldenmod_')[-1][:-3]


def _align(ours_opts, ref_opts, lp):
    perm, start = [], 0
    for mine, theirs in zip(ours_opts, ref_opts):
        perm += [start + theirs.index(m) for m in mine]
        start += len(theirs)
    return lp[perm]


def _close(x, ref, rtol=2e-5, atol=2e-6):
    x, ref = x.double().cpu(), ref.double()
    prob_ok = (x.exp() - ref.exp()).abs() <= 5e-7  # saturated entries: compared in probability space
    return bool((((x - ref).abs() <= rtol * ref.abs() + atol) | prob_ok).all())


def _check_grads(named, ref_grads, tag, rtol=1e-4):
    for k, g_mine in named:
        g = ref_grads[k].double()
        err = (g_mine.detach().cpu().double() - g).abs().max()
        assert err <= rtol * g.abs().max() + 1e-7, (tag, k, float(err), float(g.abs().max()))


@pytest.mark.parametrize('path', FILES, ids=_ids)
def test_modulated_forward_backward_matches_reference_golden(path):
    from dfol_vqa_b200.interpreter import FusedTrainStep
    case = helpers.load_golden(path)
    ont = helpers.ontology_of(case)
    ref = case['ref32']
    nets = helpers.attention_networks_of(case)
    interp = helpers.build_interpreter(ont, case['dims'], case['state'], attention_nets=nets)
    assert interp._has_modulator
    pbs = helpers.to_cuda(helpers.program_batches_of(case))

    # (1) drop-in surface + autograd (oracle parameters through our backward kernels, attention networks through
    # d loss / d modulations from the backward interpreter)
    interp.train()
    result = interp(pbs, True)
    lp = result['log_probability']
    ref_lp = ref['log_probability']
    if ref['type'] == 1 and case['terminal'] != 'compare':
        ref_lp = _align(result['options'], ref['options'], ref_lp)
    assert _close(lp.detach(), ref_lp), (lp.detach().cpu() - ref_lp).abs().max()
    answers = [a for pb in pbs for a in pb._answers]
    loss = orc.compute_loss([{'log_probability': lp, 'type': result['type'], 'options': result['options']}],
                            [answers]) / len(answers)
    loss.backward()
    assert abs(float(loss) - float(ref['loss'])) <= 2e-5 * max(1.0, abs(float(ref['loss'])))
    sd_keys = {id(p): k for k, p in interp.named_parameters()}
    _check_grads([(sd_keys[id(p)], p.grad) for p in interp.oracle_parameters()], ref['grads'], 'oracle')
    att = [(name + '.' + pn, p.grad if p.grad is not None else torch.zeros_like(p))
           for net, name in zip(nets, helpers.ATTENTION_NETS) for pn, p in net.named_parameters()]
    _check_grads(att, ref['grads'], 'attention')

    # the calibrator can be switched off per call (modulator_switch, batch_base_interpreter.py:72)
    with torch.no_grad():
        off = interp(pbs, True, modulator_switch=False)['log_probability']
    assert not torch.allclose(off, lp.detach(), rtol=1e-3, atol=1e-4)

    # (2) fused train step: same gradients in the flat bucket
    interp.zero_grad()
    step = FusedTrainStep(interp)
    loss2 = step.forward_backward(pbs)
    assert abs(float(loss2) - float(ref['loss'])) <= 2e-5 * max(1.0, abs(float(ref['loss'])))
    _check_grads([(sd_keys[id(p)], step.grads[id(p)]) for p in interp.oracle_parameters()], ref['grads'], 'fused')
    att = [(name + '.' + pn, step.grads[id(p)])
           for net, name in zip(nets, helpers.ATTENTION_NETS) for pn, p in net.named_parameters()]
    _check_grads(att, ref['grads'], 'fused attention')
    step.optimizer_step()
    torch.cuda.synchronize()
    assert all(torch.isfinite(p).all() for p in interp.parameters())


@pytest.mark.parametrize('path', FILES, ids=_ids)
def test_modulation_gradient_matches_oracle(path):
    """Kernel-level: d loss / d modulations written by the backward interpreter against autograd through the oracle's
    apply_modulations (random modulation rows, away from the identity)."""
    from dfol_vqa_b200.engine import SceneLayout
    case = helpers.load_golden(path)
    ont = helpers.ontology_of(case)
    nets = helpers.attention_networks_of(case)
    interp = helpers.build_interpreter(ont, case['dims'], case['state'], attention_nets=nets)
    host_pbs = helpers.program_batches_of(case)
    pbs = helpers.to_cuda(host_pbs)
    cp = interp.compiled(pbs[0], False)
    torch.manual_seed(3)
    rows = (0.04 + 0.12 * torch.rand(cp.mod_rows, 4))
    rows[:, 3] = 0.2 + 0.6 * torch.rand(cp.mod_rows)
    rows[:, 0] = 0.1 + 0.06 * torch.rand(cp.mod_rows)  # alpha >= 1 keeps the case well-conditioned (see below)
    # oracle side, fp32 (the restatement) and fp64 (noise floor of the reference's own fp32 formulas)
    w = None
    ref = {}
    for dtype in (torch.float32, torch.float64):
        r = rows.detach().clone().to(dtype).requires_grad_(True)
        mods = {(s, k): r[b:b + n] for s, k, n, b in cp.mod_plan}
        params = {k: v.to(dtype) for k, v in case['state'].items()}
        results, _ = orc.run_step(ont, params, helpers.program_batches_of(case, dtype), True, modulations=[mods])
        lp_ref = results[0]['log_probability']
        if w is None:
            w = torch.linspace(0.5, 1.5, lp_ref.numel())
        (lp_ref * w.to(dtype)).sum().backward()
        ref[dtype] = (lp_ref.detach(), r.grad.clone())
    lp32, g32 = ref[torch.float32]
    lp64, g64 = ref[torch.float64]
    # CUDA side
    eng = interp._engine
    layout = SceneLayout.get(case['counts'], len(ont._vocabulary['idx_to_arg']), len(ont._relation_index),
                             torch.device('cuda', 0))
    scene = eng.build_scene(pbs[0]._object_features.float().contiguous(), layout, cp=cp)
    scene.mods = rows.cuda().contiguous()
    scene.d_mods = torch.zeros_like(scene.mods)
    lp, tape = eng.run_programs(cp, scene, save_tape=True)
    ok, worst = helpers.close_to_reference(lp.cpu(), lp32, lp64, rtol=3e-5, atol=2e-6)
    assert ok, ('log_probability', worst)
    eng.program_backward(cp, scene, tape, w.cuda())
    err = (scene.d_mods.cpu().double() - g64).abs().max()
    noise = (g32.double() - g64).abs().max()
    assert err <= 1e-4 * g64.abs().max() + 4 * noise + 1e-6, (float(err), float(g64.abs().max()), float(noise))


@pytest.mark.parametrize('path', [p for p in FILES if any(t in p for t in ('verify_rel', 'query_attr', 'choose_rel'))],
                         ids=_ids)
def test_frozen_oracle_trains_attention_networks_only(path):
    """sample_config.yaml's arrangement (all four oracle networks frozen, only the attention networks train) takes the
    short backward path (backward interpreter only) and still yields the reference's attention gradients."""
    case = helpers.load_golden(path)
    ont = helpers.ontology_of(case)
    ref = case['ref32']
    pbs = helpers.to_cuda(helpers.program_batches_of(case))
    answers = [a for pb in pbs for a in pb._answers]
    nets = helpers.attention_networks_of(case)
    interp = helpers.build_interpreter(ont, case['dims'], case['state'], attention_nets=nets, freeze_oracle=True)
    interp.train()
    result = interp(pbs, True)
    loss = orc.compute_loss([{'log_probability': result['log_probability'], 'type': result['type'],
                              'options': result['options']}], [answers]) / len(answers)
    loss.backward()
    assert all(p.grad is None for p in interp.oracle_parameters())
    att = [(name + '.' + pn, p.grad if p.grad is not None else torch.zeros_like(p))
           for net, name in zip(nets, helpers.ATTENTION_NETS) for pn, p in net.named_parameters()]
    _check_grads(att, ref['grads'], 'attention (frozen oracle)')

    from dfol_vqa_b200.interpreter import FusedTrainStep
    interp.zero_grad()
    step = FusedTrainStep(interp)
    assert not step.oracle_trainable and step.flat.numel() == sum(p.numel() for p in interp.attention_parameters())
    step.forward_backward(pbs)
    att = [(name + '.' + pn, step.grads[id(p)])
           for net, name in zip(nets, helpers.ATTENTION_NETS) for pn, p in net.named_parameters()]
    _check_grads(att, ref['grads'], 'fused attention (frozen oracle)')
    before = step.flat.clone()
    step.optimizer_step()
    assert not torch.equal(before, step.flat)


@pytest.mark.parametrize('terminal,n_max,ragged', [('chain', 48, False), ('verify_rel', 37, True), ('choose_rel', 24, True),
                                                   ('query_attr', 48, False), ('two_same', 30, True),
                                                   ('compare', 100, False)])
def test_fast_interpreter_with_modulations_matches_exact(terminal, n_max, ragged):
    """Tensor-core mode: dfol_program_{fwd,bwd}_fast with modulations (bulk-async tile ring, probability-space relate)
    against the exact kernels on the same tables and the same random modulation rows, at the reference's real
    dimensions: log-probabilities, compact gradient slices and d loss / d modulations."""
    from test_gpu_tc_kernels import _programs_world
    from dfol_vqa_b200.engine import SceneLayout
    from dfol_vqa_b200.networks import build_attention_networks
    ont, dims, pbs = _programs_world(terminal, 12, n_max, ragged, seed=41)
    nets = build_attention_networks(dims['emb'], 50)
    interp = helpers.build_interpreter(ont, dims, seed=5, gemm_mode='bf16', emb_bias=-4.0,
                                       attention_nets=[nets[k] for k in ('forward_attention_network',
                                                                         'backward_attention_network',
                                                                         'attention_output_network')])
    pb = pbs[0].to_cuda(0)
    cp = interp.compiled(pb, False)
    assert cp.mod_rows > 0
    ident = interp.modulations(cp)   # reference initialisation: identity modulations
    assert ident.shape == (cp.mod_rows, 4)
    counts = interp._object_counts(pb)
    layout = SceneLayout.get(counts, interp._weights.emb.weight.shape[0], len(ont._relation_index), torch.device('cuda', 0))
    eng = interp._engine
    torch.manual_seed(9)
    rows = 0.06 + 0.08 * torch.rand(cp.mod_rows, 4, device='cuda')
    rows[:, 3] = 0.3 + 0.4 * torch.rand(cp.mod_rows, device='cuda')
    # alpha >= 1: an exponent below one stretches probabilities of ~1e-8 -- where the fp32 rounding of 1 - p decides
    # between log(6e-8) and the -46 clamp, differently for e^{l} e^{a} (fast build) and e^{l + a} (exact build and the
    # reference) -- into the 1e-4 range, and the comparison would measure that coin flip instead of the kernels
    rows[:, 0] = 0.1 + 0.04 * torch.rand(cp.mod_rows, device='cuda')
    with torch.no_grad():
        scene = eng.build_scene(pb._object_features.float(), layout, keep_for_backward=True, cp=cp)
        plain, _ = eng.run_programs(cp, scene, save_tape=False)
        scene.mods = ident.detach().float().contiguous()
        same, _ = eng.run_programs(cp, scene, save_tape=False)
        sat = (same.exp() - plain.exp()).abs() <= 1e-6
        assert bool((((same - plain).abs() <= 2e-3 * plain.abs() + 2e-4) | sat).all()), (same, plain)
        out = {}
        for mode in ('bf16', 'fp32'):   # 'bf16' -> *_fast entry points, 'fp32' -> exact entry points
            eng.gemm_mode = mode
            scene.mods = rows.contiguous()
            scene.d_mods = torch.zeros_like(rows)
            lp, tape = eng.run_programs(cp, scene, save_tape=True)
            d_lp = torch.linspace(-1.0, 1.0, lp.numel(), device='cuda')
            g_attr, g_rel = eng.program_backward(cp, scene, tape, d_lp)
            out[mode] = (lp.clone(), g_attr.clone(), g_rel.clone(), scene.d_mods.clone())
        eng.gemm_mode = 'bf16'
    lp_f, lp_e = out['bf16'][0], out['fp32'][0]
    ok = (lp_f - lp_e).abs() <= 2e-3 * lp_e.abs() + 2e-4
    sat = (lp_f.exp() - lp_e.exp()).abs() <= 1e-6
    assert bool((ok | sat).all()), (lp_f, lp_e)
    assert not torch.allclose(lp_e, plain, rtol=1e-3, atol=1e-4)
    for a, b in zip(out['bf16'][1:], out['fp32'][1:]):
        scale = float(b.abs().max())
        if scale < 1e3:
            assert float((a - b).abs().max()) <= 2e-2 * scale + 1e-6, float((a - b).abs().max()) / (scale + 1e-9)


@pytest.mark.parametrize('terminal,n_max', [('verify_rel', 20), ('query_attr', 12), ('choose_rel', 16),
                                            ('two_same', 10), ('compare', 14), ('and', 18)])
@pytest.mark.parametrize('tape', [True, False])
def test_native_modulator_matches_torch_statement(terminal, n_max, tape, monkeypatch):
    """modulator_cuda.NativeAttentionTransfer (tape: the two persistent kernels of csrc/modulator_tape.cu; not tape: one
    launch per step; hand-written LSTM-cell / output-layer kernels, hand-derived backward)
    against modulator.AttentionTransfer (torch ops + autograd; held to the recorded reference runs by the CPU tests) at
    the reference's real dimensions (318-wide features, state 50): modulation rows and all 10 parameter gradients."""
    from test_gpu_tc_kernels import _programs_world
    from dfol_vqa_b200.compiler import ProgramCompiler
    from dfol_vqa_b200.modulator import AttentionTransfer
    from dfol_vqa_b200.modulator_cuda import NativeAttentionTransfer
    from dfol_vqa_b200.networks import build_attention_networks
    from dfol_vqa_b200 import modulator_cuda
    monkeypatch.setattr(modulator_cuda, '_USE_TAPE', tape)
    ont, dims, pbs = _programs_world(terminal, 24, n_max, True, seed=77)
    torch.manual_seed(11)
    nets = build_attention_networks(dims['emb'], 50)
    with torch.no_grad():
        nets['attention_output_network'][0].weight.normal_(0.0, 0.3)
    fwd, bwd, out = (nets[k].cuda() for k in ('forward_attention_network', 'backward_attention_network',
                                              'attention_output_network'))
    cp = ProgramCompiler(ont, normalize=True, modulated=True).compile(pbs[0], [n_max] * 24)
    ref = AttentionTransfer(fwd, bwd, out, ont)
    rows = ref.modulations(cp)
    # rows of questions that do not execute a slot (mask 0) are computed by the reference but never read: the
    # interpreter gates those questions back to their input (batch_base_interpreter.py:166-167), no instruction
    # carries their row, their d_mods is zero.  The native pass does not reproduce them (it keeps the gated state).
    live = torch.ones(rows.shape[0], dtype=torch.bool)
    for slot, key, n, base in cp.mod_plan:
        m = cp.mod_descs[slot]['mask']
        if m is not None and n == len(m):
            live[base:base + n] = torch.tensor(m) > 0
    live = live.cuda()
    assert float(live.float().mean()) < 1.0
    w = torch.randn(rows.shape, device='cuda', generator=torch.Generator('cuda').manual_seed(5)) * live[:, None]
    rows.backward(w)
    params = ref.parameters()
    ref_grads = [p.grad.clone() for p in params]
    native = NativeAttentionTransfer(fwd, bwd, out, ont)
    mods, ctx = native.forward(cp)
    assert mods.shape == rows.shape
    assert torch.allclose(mods[live], rows.detach()[live], rtol=1e-5, atol=1e-6), (mods - rows.detach()).abs().max()
    grads = {id(p): torch.zeros_like(p) for p in params}
    native.backward(ctx, w, grads)
    for p, g in zip(params, ref_grads):
        err = (grads[id(p)] - g).abs().max()
        assert err <= 2e-4 * g.abs().max() + 1e-6, (tuple(p.shape), float(err), float(g.abs().max()))


def test_full_size_calibrator_question_independence_and_additivity():
    """Size-independent properties of the calibrator arrangement at the full c1 size (B=256, N=48, real dimensions,
    frozen oracle, randomised attention networks): a question's log-probability does not depend on which other
    questions share its batch -- although the slot alignment, the masks and therefore every LSTM launch differ -- and
    the attention-network gradients of the whole batch equal the accumulation over two sub-batches."""
    import json
    from dfol_vqa_b200 import synth
    from dfol_vqa_b200.interpreter import FusedTrainStep
    from dfol_vqa_b200.networks import build_attention_networks
    from dfol_vqa_b200.ontology import synthetic_ontology
    from dfol_vqa_b200.programs import ProgramCollater
    dims = dict(box=2048, feat=512, hidden=256, emb=300)
    batch, n, sub = 256, 48, 16
    ont = synthetic_ontology(seed=1, embedding_dim=300, concept_num=2335, relation_num=333, category_num=31, class_num=53)
    torch.manual_seed(21)
    nets = build_attention_networks(300, 50)
    with torch.no_grad():
        nets['attention_output_network'][0].weight.normal_(0.0, 0.1)
    interp = helpers.build_interpreter(ont, dims, seed=0, gemm_mode='bf16', emb_bias=-4.0, freeze_oracle=True,
                                       attention_nets=[nets[k] for k in ('forward_attention_network',
                                                                         'backward_attention_network',
                                                                         'attention_output_network')])
    questions = synth.make_questions(ont, batch, 'verify_rel', 1, 3, seed=15, relate_prob=0.35)
    feats, bidx = synth.make_object_features([n] * batch, 2048, seed=16)
    full = ProgramCollater(1, lambda qs: (feats, bidx)).collate(json.loads(json.dumps(questions)))
    part = ProgramCollater(1, lambda qs: (feats[:sub * n].clone(), bidx[:sub * n].clone())).collate(
        json.loads(json.dumps(questions[:sub])))
    halves = ProgramCollater(2, helpers.slicing_source(feats, bidx)).collate(json.loads(json.dumps(questions)))
    interp.train()
    with torch.no_grad():
        lp_full = interp(helpers.to_cuda(full), True)['log_probability'].clone()
        lp_part = interp(helpers.to_cuda(part), True)['log_probability'].clone()
        lp_off = interp(helpers.to_cuda(full), True, modulator_switch=False)['log_probability'].clone()
    assert bool(torch.isfinite(lp_full).all()) and not torch.allclose(lp_full, lp_off, rtol=1e-3, atol=1e-4)
    # One batch-composition dependence is the reference's own: a blank predicate ('_' name of a relate, say) is
    # modulated iff SOME question of the batch has a predicate in that slot (FilterBatch returns early otherwise,
    # batch_base_ops.py:315-317).  Questions whose modulated sub-operators differ between the two batches are excluded.
    cf, cpart = interp.compiled(helpers.to_cuda(full)[0], False), interp.compiled(helpers.to_cuda(part)[0], False)

    def coverage(cp, q):
        rows = cp.instr[cp.q_instr[q]:cp.q_instr[q + 1]]
        return [(int(r[0]), int(r[9]) >= 0, int(r[10]) >= 0) for r in rows]
    same = torch.tensor([coverage(cf, q) == coverage(cpart, q) for q in range(sub)], device='cuda')
    assert int(same.sum()) >= sub - 4
    sat = (lp_full[:sub].exp() - lp_part.exp()).abs() <= 1e-6
    ok = ((lp_full[:sub] - lp_part).abs() <= 1e-4 * lp_part.abs() + 1e-5) | sat
    assert bool(ok[same].all()), (lp_full[:sub] - lp_part).abs()
    step = FusedTrainStep(interp)
    loss_full = float(step.forward_backward(helpers.to_cuda(full), global_question_num=batch))
    g_full = step.flat_grad.clone()
    loss_split = float(step.forward_backward(helpers.to_cuda(halves), global_question_num=batch))
    g_split = step.flat_grad.clone()
    assert abs(loss_full - loss_split) <= 1e-4 * max(1.0, abs(loss_full))
    scale = float(g_full.abs().max())
    assert scale > 0 and bool(torch.isfinite(g_full).all())
    assert float((g_full - g_split).abs().max()) <= 2e-3 * scale, float((g_full - g_split).abs().max()) / scale


def test_modulated_eval_answers_match_reference_golden():
    for path in FILES:
        case = helpers.load_golden(path)
        ont = helpers.ontology_of(case)
        interp = helpers.build_interpreter(ont, case['dims'], case['state'],
                                           attention_nets=helpers.attention_networks_of(case))
        pbs = helpers.to_cuda(helpers.program_batches_of(case))
        interp.eval()
        with torch.no_grad():
            result = interp(pbs, False)
        assert [sorted(a) for a in result['answer']] == [sorted(a) for a in case['ref32']['answer']], path


def test_state_dict_keys_follow_the_reference():
    case = helpers.load_golden(FILES[0])
    ont = helpers.ontology_of(case)
    interp = helpers.build_interpreter(ont, case['dims'], case['state'],
                                       attention_nets=helpers.attention_networks_of(case))
    keys = set(interp.state_dict().keys())
    for k in ('_ops.select._filter._forward_attention_network.weight_ih',
              '_ops.relate._relate._backward_attention_network.bias_hh',
              '_ops.verify_rel._gqa_relate._gqa_select._filter._attention_output_network.0.weight',
              '_ops.two_different._gqa_two_same._filter._attention_output_network.0.bias'):
        assert k in keys, k
    named = dict(interp.named_parameters())
    assert '_ops.select._filter._forward_attention_network.weight_ih' in named
