"""Parity of the CUDA path (through the C ABI) with the reference goldens and with the CPU oracle. Needs a GPU."""

import json

import numpy as np
import pytest
import torch

import helpers
import dfol_oracle as orc

pytestmark = pytest.mark.gpu


def _ids(p):
    return p.split('golden_')[-1][:-3]


def _align(ours_opts, ref_opts, lp):
    perm, start = [], 0
    for mine, theirs in zip(ours_opts, ref_opts):
        perm += [start + theirs.index(m) for m in mine]
        start += len(theirs)
    return lp[perm]


@pytest.mark.parametrize('path', helpers.golden_files(), ids=_ids)
def test_forward_backward_matches_reference_golden(path):
    """fp32 mode: log-probabilities, loss and all 12 parameter gradients against the recorded reference run."""
    from dfol_vqa_b200.interpreter import FusedTrainStep
    case = helpers.load_golden(path)
    ont = helpers.ontology_of(case)
    interp = helpers.build_interpreter(ont, case['dims'], case['state'])
    pbs = helpers.to_cuda(helpers.program_batches_of(case))
    ref32, ref64 = case['ref32'], case['ref64']

    # (1) drop-in surface: forward in training mode + torch autograd through the custom Function
    interp.train()
    result = interp(pbs, True)
    lp = result['log_probability']
    assert result['type'] == ref32['type']
    ref_lp, ref64_lp = ref32['log_probability'], ref64['log_probability']
    if ref32['type'] == 1 and case['terminal'] != 'compare':
        assert [sorted(o) for o in result['options']] == [sorted(o) for o in ref32['options']]
        ref_lp = _align(result['options'], ref32['options'], ref_lp)
        ref64_lp = _align(result['options'], ref32['options'], ref64_lp)
    ok, worst = helpers.close_to_reference(lp.detach().cpu(), ref_lp, ref64_lp)
    assert ok, ('log_probability', worst)

    # loss exactly as the reference trainer computes it, then autograd into our backward kernels
    answers = [a for pb in pbs for a in pb._answers]
    res_cpu_like = {'log_probability': lp, 'type': result['type'], 'options': result['options']}
    loss = orc.compute_loss([res_cpu_like], [answers]) / len(answers)
    loss.backward()
    assert abs(float(loss) - float(ref32['loss'])) <= 1e-5 * max(1.0, abs(float(ref32['loss']))) + \
        4 * abs(float(ref32['loss']) - float(ref64['loss']))
    sd_keys = {id(p): k for k, p in interp.named_parameters()}
    checked = 0
    for p in interp.oracle_parameters():
        k = sd_keys[id(p)]
        g32, g64 = ref32['grads'][k], ref64['grads'][k]
        scale = g64.abs().max().clamp(min=1e-12)
        err = (p.grad.detach().cpu().double() - g32.double()).abs().max()
        noise = (g32.double() - g64).abs().max()
        assert err <= 1e-5 * scale + 4 * noise + 1e-9, (k, float(err), float(scale), float(noise))
        checked += 1
    assert checked == 12

    # (2) fused train step: same gradients in the flat bucket, loss from the loss kernel
    interp.zero_grad()
    step = FusedTrainStep(interp)
    loss2 = step.forward_backward(pbs)
    assert abs(float(loss2) - float(ref32['loss'])) <= 1e-5 * max(1.0, abs(float(ref32['loss']))) + \
        4 * abs(float(ref32['loss']) - float(ref64['loss']))
    for p in interp.oracle_parameters():
        k = sd_keys[id(p)]
        g32, g64 = ref32['grads'][k], ref64['grads'][k]
        scale = g64.abs().max().clamp(min=1e-12)
        err = (step.grads[id(p)].cpu().double() - g32.double()).abs().max()
        noise = (g32.double() - g64).abs().max()
        assert err <= 1e-5 * scale + 4 * noise + 1e-9, ('fused', k, float(err), float(scale))


@pytest.mark.parametrize('path', helpers.golden_files(), ids=_ids)
def test_eval_answers_match_reference_golden(path):
    case = helpers.load_golden(path)
    ont = helpers.ontology_of(case)
    interp = helpers.build_interpreter(ont, case['dims'], case['state'])
    pbs = helpers.to_cuda(helpers.program_batches_of(case))
    interp.eval()
    with torch.no_grad():
        result = interp(pbs, False)
    assert [sorted(a) for a in result['answer']] == [sorted(a) for a in case['ref32']['answer']]


def test_scene_tables_match_reference_golden():
    """Attribute / relation tables of the fused scene build against the reference's compute_all_log_likelihood_2."""
    from dfol_vqa_b200.engine import SceneLayout
    path = [p for p in helpers.golden_files() if 'verify_rel' in p][0]
    case = helpers.load_golden(path)
    ont = helpers.ontology_of(case)
    interp = helpers.build_interpreter(ont, case['dims'], case['state'])
    counts = case['counts']
    C, nR = len(ont._vocabulary['idx_to_arg']), len(ont._relation_index)
    layout = SceneLayout.get(counts, C, nR, torch.device('cuda', 0))
    scene = interp._engine.build_scene(case['features'].cuda(), layout)
    ref = case['ref32']['scene']
    attr = scene.attr_ll.cpu()
    a_blk, a_str = layout.attr_blk.cpu(), layout.attr_stride.cpu()
    rows = []
    for b, n in enumerate(counts):
        blk = attr[int(a_blk[b]):int(a_blk[b]) + C * int(a_str[b])].view(C, int(a_str[b]))
        rows.append(blk[:, :n].t())
    assert torch.allclose(torch.cat(rows), ref['attr'], rtol=1e-5, atol=1e-6)
    rel = scene.rel_ll.cpu()
    r_blk, r_str = layout.rel_blk.cpu(), layout.rel_stride.cpu()
    img, s, o = ref['index']
    starts = np.concatenate([[0], np.cumsum(counts)])
    mine = []
    for b, si, oi in zip(img.tolist(), s.tolist(), o.tolist()):
        n = counts[b]
        blk = rel[int(r_blk[b]):int(r_blk[b]) + nR * int(r_str[b])].view(nR, int(r_str[b]))
        mine.append(blk[:, (si - starts[b]) * n + (oi - starts[b])])
    assert torch.allclose(torch.stack(mine), ref['rel'], rtol=1e-5, atol=1e-6)
    for b, n in enumerate(counts):
        blk = rel[int(r_blk[b]):int(r_blk[b]) + nR * int(r_str[b])].view(nR, int(r_str[b]))
        assert bool((blk[:, [i * n + i for i in range(n)]] == -30.0).all())


def test_inner_module_surfaces_match_reference_golden():
    """The inner surfaces of the three modules -- FastBoxFeaturizer.featurize_scene (reference
    batch_gqa_boxfeatures_pipeline.py:199-281), FastClassifierOracle.compute_all_log_likelihood_2
    (classifier_oracle.py:145-156), FastGQAInterpreter.build_scene (batch_base_interpreter.py:45-70) -- against the values
    the UNMODIFIED reference returned from the same calls (recorded in the golden's ``scene`` entry)."""
    path = [p for p in helpers.golden_files() if 'verify_rel' in p][0]
    case = helpers.load_golden(path)
    ont = helpers.ontology_of(case)
    interp = helpers.build_interpreter(ont, case['dims'], case['state'])
    ref = case['ref32']['scene']
    dev = torch.device('cuda', 0)
    feats, bidx = case['features'], case['batch_index']
    out = interp._featurizer.featurize_scene(dev, feats, bidx, {})
    assert out['object_num'] == feats.shape[0]
    for mine, theirs in zip(out['relation_features']['index'], ref['index']):
        assert torch.equal(mine.cpu(), theirs.long())
    # the feature rows themselves against the oracle's restatement of the same function
    F = case['dims']['feat']
    f = torch.sigmoid(torch.nn.functional.linear(feats[:, :-6], case['state']['_featurizer._featurizer_network._network.1.weight'],
                                                 case['state']['_featurizer._featurizer_network._network.1.bias']))
    assert torch.allclose(out['attribute_features'][:, :F].cpu(), f, rtol=1e-5, atol=1e-6)
    img, s, o = (t.long() for t in ref['index'])
    objs = out['attribute_features'].cpu()
    rel = out['relation_features']['features'].cpu()
    assert rel.shape == (img.numel(), 2 * (F + 4) + 4)
    assert torch.equal(rel[:, :F + 4], objs[s]) and torch.equal(rel[:, F + 4:2 * (F + 4)], objs[o])
    pos = objs[:, F:]
    dy = pos[s, 1] + pos[s, 3] / 2 - pos[o, 1] - pos[o, 3] / 2
    dist = torch.sqrt((pos[s, 0] + pos[s, 2] / 2 - pos[o, 0] - pos[o, 2] / 2) ** 2 + dy ** 2)
    geo = torch.stack([dist, torch.asin(dy / dist.clamp(min=1e-10)), (pos[o, 0] - pos[s, 0]).sign(),
                       (pos[o, 1] - pos[s, 1]).sign()], dim=1)
    assert torch.allclose(rel[:, 2 * (F + 4):], geo, rtol=1e-5, atol=1e-6)
    attr_ll, rel_ll = interp._oracle.compute_all_log_likelihood_2(out['attribute_features'],
                                                                  out['relation_features']['features'])
    assert torch.allclose(attr_ll.cpu(), ref['attr'], rtol=1e-5, atol=1e-6)
    assert torch.allclose(rel_ll.cpu(), ref['rel'], rtol=1e-5, atol=1e-6)
    scene = interp.build_scene(dev, feats, bidx, {})
    from dfol_vqa_b200.engine import SceneLayout
    C, nR = len(ont._vocabulary['idx_to_arg']), len(ont._relation_index)
    layout = SceneLayout.get(case['counts'], C, nR, dev)
    direct = interp._engine.build_scene(feats.cuda(), layout)
    a_blk, a_str = layout.attr_blk.cpu(), layout.attr_stride.cpu()
    r_blk, r_str = layout.rel_blk.cpu(), layout.rel_stride.cpu()
    for b, n in enumerate(case['counts']):   # (the padding between table slices is uninitialised: valid entries only)
        for mine, theirs, blk, stride, rows, width in ((scene.attr_ll, direct.attr_ll, a_blk, a_str, C, n),
                                                       (scene.rel_ll, direct.rel_ll, r_blk, r_str, nR, n * n)):
            lo, ld = int(blk[b]), int(stride[b])
            assert torch.equal(mine[lo:lo + rows * ld].view(rows, ld)[:, :width],
                               theirs[lo:lo + rows * ld].view(rows, ld)[:, :width])
    with pytest.raises(AssertionError):
        interp._oracle.compute_all_log_likelihood_2(out['attribute_features'].cpu(), None)   # no CPU fallback


@pytest.mark.parametrize('terminal,batch,n_max,ragged', [
    ('exist', 16, 48, False), ('verify_rel', 12, 37, True), ('query_attr', 8, 24, True), ('choose_rel', 8, 100, False),
    ('and', 10, 48, True), ('two_same', 6, 31, True), ('compare', 6, 20, True), ('all_same', 6, 17, True)])
def test_cuda_matches_oracle_at_reference_dims(terminal, batch, n_max, ragged):
    """Full-size oracle dims (2048 -> 512 -> 256 -> 300 -> C) on a 400-concept synthetic vocabulary, up to N = 100:
    fp32 CUDA path vs the CPU oracle run in fp32 and fp64 on the same seeded inputs (forward + all gradients)."""
    from dfol_vqa_b200 import synth
    from dfol_vqa_b200.ontology import synthetic_ontology
    from dfol_vqa_b200.programs import ProgramCollater
    from dfol_vqa_b200.interpreter import FusedTrainStep
    dims = dict(box=2048, feat=512, hidden=256, emb=300)
    ont = synthetic_ontology(400, 60, 6, 5, seed=3, embedding_dim=300)
    interp = helpers.build_interpreter(ont, dims, seed=5)
    questions = synth.make_questions(ont, batch, terminal, 1, 5, seed=7)
    counts = synth.object_counts(batch, n_max, ragged, seed=9)
    feats, bidx = synth.make_object_features(counts, 2048, seed=11)
    pbs = ProgramCollater(1, lambda qs: (feats, bidx)).collate(questions)

    outs = {}
    for dtype in (torch.float32, torch.float64):
        params = helpers.oracle_params(interp, dtype, requires_grad=True)
        pb_cpu = ProgramCollater(1, lambda qs: (feats.to(dtype), bidx)).collate(json.loads(json.dumps(questions)))
        results, loss = orc.run_step(ont, params, pb_cpu, is_training=True)
        loss.backward()
        outs[dtype] = (results[0]['log_probability'].detach(), loss.detach(), {k: v.grad for k, v in params.items()})

    step = FusedTrainStep(interp)
    loss = step.forward_backward(helpers.to_cuda(pbs))
    interp.eval()
    result = interp(helpers.to_cuda(pbs), True)
    lp32, l32, g32 = outs[torch.float32]
    lp64, l64, g64 = outs[torch.float64]
    ok, worst = helpers.close_to_reference(result['log_probability'].detach().cpu(), lp32, lp64)
    assert ok, worst
    assert abs(float(loss) - float(l32)) <= 1e-5 * max(1.0, abs(float(l32))) + 4 * abs(float(l32) - float(l64))
    keys = {id(p): k for k, p in interp.named_parameters()}
    for p in interp.oracle_parameters():
        k = keys[id(p)]
        scale = g64[k].abs().max().clamp(min=1e-12)
        err = (step.grads[id(p)].cpu().double() - g32[k].double()).abs().max()
        noise = (g32[k].double() - g64[k]).abs().max()
        assert err <= 1e-5 * scale + 4 * noise + 1e-9, (k, float(err), float(scale), float(noise))


@pytest.mark.parametrize('terminal', ['exist', 'and', 'query_attr', 'verify_rel'])
def test_return_trace_matches_oracle(terminal):
    """return_trace=True (what VQATrainer._visualize_batch consumes, trainer.py:548-591): the per-slot log-attention
    rebuilt from the attention tape against the oracle's per-image trace."""
    path = [p for p in helpers.golden_files() if ('golden_%s_s1' % terminal) in p][0]
    case = helpers.load_golden(path)
    ont = helpers.ontology_of(case)
    interp = helpers.build_interpreter(ont, case['dims'], case['state'])
    host = helpers.program_batches_of(case)
    pbs = helpers.to_cuda(host)
    interp.eval()
    with torch.no_grad():
        result, traces = interp(pbs, False, return_trace=True)
    params = {k: v.clone() for k, v in case['state'].items()}
    ref = orc.OracleInterpreter(ont, params).run(host[0], is_training=False)
    starts = np.concatenate([[0], np.cumsum(case['counts'])])
    slots = [t for t in ref['trace'] if t is not None]
    assert len(traces) == 1 and len(traces[0]) == len(slots) >= 1
    for entry, (atts, names) in zip(traces[0], slots):
        assert entry._log_attention.shape == (len(case['counts']), int(starts[-1]))
        assert list(entry._name) == list(names)
        for q, a in enumerate(atts):
            mine = entry._log_attention[q, starts[q]:starts[q + 1]].cpu()
            assert torch.allclose(mine, a, rtol=1e-5, atol=1e-5), (q, mine, a)


@pytest.mark.parametrize('terminal', ['exist', 'and', 'verify_rel', 'query_attr', 'choose_rel', 'two_same', 'all_same',
                                      'compare', 'all_different', 'two_different', 'verify_attrs', 'choose_attr', 'or'])
def test_hard_mode_eval_matches_oracle(terminal):
    """hard_mode (config `hard_mode: True`; min instead of sum in the quantifiers when answers are given,
    batch_base_types.py:104-112): eval-mode log-probabilities and answers against the oracle; training-mode passes of the
    same interpreter stay soft."""
    path = [p for p in helpers.golden_files() if ('golden_%s_s1' % terminal) in p][0]
    case = helpers.load_golden(path)
    ont = helpers.ontology_of(case)
    interp = helpers.build_interpreter(ont, case['dims'], case['state'], hard_mode=True)
    host = helpers.program_batches_of(case)
    pbs = helpers.to_cuda(host)
    interp.eval()
    with torch.no_grad():
        result = interp(pbs, False)
        soft = interp(pbs, True)['log_probability'].cpu()
    params = {k: v.clone() for k, v in case['state'].items()}
    ref = orc.OracleInterpreter(ont, params, hard_mode=True).run(host[0], is_training=False)
    ref_lp = ref['log_probability']
    lp = result['log_probability'].cpu()
    if ref['type'] == 1 and terminal != 'compare':
        ref_lp = _align(result['options'], ref['options'], ref_lp)
    sat = (lp.exp() - ref_lp.exp()).abs() <= 5e-7
    assert bool((((lp - ref_lp).abs() <= 2e-5 * ref_lp.abs() + 2e-6) | sat).all()), (lp, ref_lp)
    assert [sorted(a) for a in result['answer']] == [sorted(a) for a in ref['answer']]
    import os
    rec = torch.load(os.path.join(helpers.GOLDEN_DIR, 'hard_eval_golden.pt'), weights_only=False)[os.path.basename(path)]
    assert [sorted(a) for a in result['answer']] == [sorted(a) for a in rec['answer']]   # the reference's own run
    ok, _ = helpers.close_to_reference(soft, case['ref32']['log_probability'] if ref['type'] != 1 or terminal == 'compare'
                                       else _align(result['options'], case['ref32']['options'],
                                                   case['ref32']['log_probability']),
                                       case['ref64']['log_probability'] if ref['type'] != 1 or terminal == 'compare'
                                       else _align(result['options'], case['ref32']['options'],
                                                   case['ref64']['log_probability']))
    assert ok
    if terminal in ('exist', 'verify_rel', 'choose_attr'):
        assert not torch.allclose(lp, soft, rtol=1e-4, atol=1e-5)   # hard and soft quantifiers really differ
    if terminal in ('query_attr', 'all_different', 'two_different'):
        assert torch.allclose(lp, soft, rtol=1e-6, atol=1e-7)       # ... except where the reference drops hard_mode


@pytest.mark.parametrize('terminal', ['choose_attr', 'query_attr', 'choose_rel', 'all_same', 'two_same'])
def test_unnormalised_oracle_and_likelihood_threshold(terminal):
    """`normalize_oracle: False` (no softmax over a question's options, classifier_oracle.py:22-42) and a non-zero
    `likelihood_threshold` (answers need p > threshold, util.find_max_ind): training-mode log-probabilities + gradients
    and eval answers against the oracle (which a live run holds to the reference with the same settings)."""
    path = [p for p in helpers.golden_files() if ('golden_%s_s1' % terminal) in p][0]
    case = helpers.load_golden(path)
    ont = helpers.ontology_of(case)
    interp = helpers.build_interpreter(ont, case['dims'], case['state'], normalize=False, likelihood_threshold=0.3)
    host = helpers.program_batches_of(case)
    pbs = helpers.to_cuda(host)
    answers = [a for pb in pbs for a in pb._answers]
    params = {k: v.clone().requires_grad_(True) for k, v in case['state'].items()}
    oi = orc.OracleInterpreter(ont, params, normalize=False, likelihood_threshold=0.3)
    ref = oi.run(host[0], is_training=True)
    interp.train()
    result = interp(pbs, True)
    lp, ref_lp = result['log_probability'], ref['log_probability']
    if ref['type'] == 1:
        ref_lp = _align(result['options'], ref['options'], ref_lp)
    sat = (lp.detach().cpu().exp() - ref_lp.detach().exp()).abs() <= 5e-7
    assert bool((((lp.detach().cpu() - ref_lp.detach()).abs() <= 2e-5 * ref_lp.detach().abs() + 2e-6) | sat).all())
    unnorm_differs = not torch.allclose(lp.detach().cpu(), case['ref32']['log_probability'] if ref['type'] != 1 else
                                        _align(result['options'], case['ref32']['options'],
                                               case['ref32']['log_probability']), rtol=1e-3, atol=1e-4)
    assert unnorm_differs or terminal == 'two_same'
    loss = orc.compute_loss([{'log_probability': lp, 'type': result['type'], 'options': result['options']}],
                            [answers]) / len(answers)
    loss.backward()
    (orc.compute_loss([ref], [answers]) / len(answers)).backward()
    keys = {id(p): k for k, p in interp.named_parameters()}
    for p in interp.oracle_parameters():
        g = params[keys[id(p)]].grad
        g = g if g is not None else torch.zeros_like(params[keys[id(p)]])
        assert (p.grad.cpu() - g).abs().max() <= 2e-4 * g.abs().max() + 1e-7, keys[id(p)]
    interp.eval()
    with torch.no_grad():
        ev = interp(pbs, False)
        ref_ev = oi.run(host[0], is_training=False)
    assert [sorted(a) for a in ev['answer']] == [sorted(a) for a in ref_ev['answer']]


@pytest.mark.parametrize('terminal', ['verify_rel', 'choose_rel', 'query_attr', 'and'])
def test_edge_object_counts(terminal):
    """Edge cases of the per-image layout: images with ONE object (a relate has no other object to quantify over: the
    empty sum gives log(1 - e^0) -> the 1e-20 clamp), two objects, and the maximum the kernels support (128), in one
    ragged batch; fp32 path (exact interpreter build) and tensor-core path (fast build) against the oracle."""
    from dfol_vqa_b200 import synth
    from dfol_vqa_b200.ontology import synthetic_ontology
    from dfol_vqa_b200.programs import ProgramCollater
    dims = dict(box=2048, feat=512, hidden=256, emb=300)
    ont = synthetic_ontology(400, 60, 6, 5, seed=3, embedding_dim=300)
    counts = [1, 2, 128, 3, 1, 127, 4, 5]
    questions = synth.make_questions(ont, len(counts), terminal, 1, 4, seed=17, relate_prob=0.7)
    feats, bidx = synth.make_object_features(counts, 2048, seed=19)
    outs = {}
    for dtype in (torch.float32, torch.float64):
        interp = helpers.build_interpreter(ont, dims, seed=5, emb_bias=-4.0)
        params = helpers.oracle_params(interp, dtype)
        pb_cpu = ProgramCollater(1, lambda qs: (feats.to(dtype), bidx)).collate(json.loads(json.dumps(questions)))
        with torch.no_grad():
            outs[dtype] = orc.OracleInterpreter(ont, params).run(pb_cpu[0], is_training=True)
    lp32, lp64 = outs[torch.float32]['log_probability'], outs[torch.float64]['log_probability']
    assert bool(torch.isfinite(lp32).all())
    for mode in ('fp32', 'bf16'):
        interp = helpers.build_interpreter(ont, dims, seed=5, gemm_mode=mode, emb_bias=-4.0)
        pbs = ProgramCollater(1, lambda qs: (feats, bidx)).collate(json.loads(json.dumps(questions)))
        interp.train()
        with torch.no_grad():
            result = interp(helpers.to_cuda(pbs), True)
        lp = result['log_probability'].cpu()
        ref32, ref64 = lp32, lp64
        if outs[torch.float32]['type'] == 1:
            ref32 = _align(result['options'], outs[torch.float32]['options'], lp32)
            ref64 = _align(result['options'], outs[torch.float32]['options'], lp64)
        if mode == 'fp32':
            ok, worst = helpers.close_to_reference(lp, ref32, ref64)
            assert ok, (mode, worst)
        else:
            # (probabilities of ~1e-6: fp32 log(1 - e^x) is only good to a few 1e-6 absolute there, SURVEY.md §7 -- the
            # fp32 and fp64 oracle runs themselves differ by that much -- so either reference may be matched)
            sat = ((lp.exp() - ref64.float().exp()).abs() <= 5e-6) | ((lp.exp() - ref32.exp()).abs() <= 1e-6)
            good = ((lp - ref64.float()).abs() <= 2e-2 * ref64.float().abs() + 2e-2) | \
                ((lp - ref32).abs() <= 2e-2 * ref32.abs() + 2e-2) | sat
            bad = (~good).nonzero().flatten().tolist()
            assert not bad, [(i, float(lp[i]), float(ref64[i]), float(ref32[i])) for i in bad[:8]]
