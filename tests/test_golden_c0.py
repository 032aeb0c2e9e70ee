"""c0-scale fixture (B = 32, N = 48, 2048-d features, 2335 concepts) recorded from the unmodified reference
(tests/golden/make_golden_c0.py): the CPU oracle (not gpu) and the CUDA path in both precision modes (gpu) are held to
it.  Inputs and initial weights are regenerated from the fixture's seeds by ``make_golden_c0.case_inputs``."""

import json
import os
import sys

import pytest
import torch

import helpers
import dfol_oracle as orc

sys.path.insert(0, helpers.GOLDEN_DIR)
import make_golden_c0 as mk  # noqa: E402

PATH = os.path.join(helpers.GOLDEN_DIR, 'goldenc0_exist.pt')


def _world():
    case = torch.load(PATH, weights_only=False)
    md, ont, questions, feats, bidx, state = mk.case_inputs(seed=case['seed'], questions=json.loads(case['questions']))
    return case, ont, questions, feats, bidx, state


def _collate(questions, feats, bidx):
    from dfol_vqa_b200.programs import ProgramCollater
    return ProgramCollater(1, lambda qs: (feats, bidx)).collate(json.loads(json.dumps(questions)))


def _check_grads(case, get, rtol, use_noise=True):
    r32, r64 = case['ref32']['grads'], case['ref64']['grads']
    for k in r32:
        g = get(k).detach().reshape(-1).double().cpu()
        idx = SAMPLE_POS[k]
        s32, s64 = r32[k]['samples'].double(), r64[k]['samples'].double()
        scale = r64[k]['max']
        noise = float((s32 - s64).abs().max()) if use_noise else 0.0
        err = float((g[idx] - s32).abs().max())
        assert err <= rtol * scale + 4 * noise + 1e-9, (k, err, scale, noise)
        assert abs(float(g.norm()) - r32[k]['norm']) <= (rtol * 4 + 1e-6) * r64[k]['norm'] + 4 * abs(
            r32[k]['norm'] - r64[k]['norm']), (k, float(g.norm()), r32[k]['norm'])


SAMPLE_POS = None


def _positions(state):
    global SAMPLE_POS
    SAMPLE_POS = mk.sample_positions(state)


def test_oracle_matches_reference_at_c0_scale():
    case, ont, questions, feats, bidx, state = _world()
    _positions(state)
    params = {k: v.clone().requires_grad_(True) for k, v in state.items()}
    res, loss = orc.run_step(ont, params, _collate(questions, feats, bidx), is_training=True)
    loss.backward()
    ref32, ref64 = case['ref32'], case['ref64']
    ok, worst = helpers.close_to_reference(res[0]['log_probability'].detach(), ref32['log_probability'],
                                           ref64['log_probability'])
    assert ok, worst
    assert abs(float(loss) - float(ref32['loss'])) <= 1e-5 * abs(float(ref32['loss'])) + 4 * abs(
        float(ref32['loss']) - float(ref64['loss']))
    _check_grads(case, lambda k: params[k].grad, 1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize('mode', ['fp32', 'bf16'])
def test_cuda_matches_reference_at_c0_scale(mode):
    from dfol_vqa_b200.interpreter import FusedTrainStep
    case, ont, questions, feats, bidx, state = _world()
    _positions(state)
    interp = helpers.build_interpreter(ont, case['dims'], state, gemm_mode=mode)
    pbs = helpers.to_cuda(_collate(questions, feats, bidx))
    ref32, ref64 = case['ref32'], case['ref64']
    interp.train()
    with torch.no_grad():
        lp = interp(pbs, True)['log_probability'].cpu()
    if mode == 'fp32':
        ok, worst = helpers.close_to_reference(lp, ref32['log_probability'], ref64['log_probability'])
        assert ok, worst
    else:
        err = (lp - ref32['log_probability']).abs()
        assert bool((err <= 2e-2 * ref32['log_probability'].abs().clamp(min=1.0)).all()), float(err.max())
    step = FusedTrainStep(interp)
    loss = float(step.forward_backward(pbs).detach())
    tol = 1e-5 if mode == 'fp32' else 2e-2
    assert abs(loss - float(ref32['loss'])) <= tol * max(1.0, abs(float(ref32['loss']))) + 4 * abs(
        float(ref32['loss']) - float(ref64['loss']))
    keys = {k: p for k, p in interp.named_parameters()}
    _check_grads(case, lambda k: step.grads[id(keys[k])], 1e-5 if mode == 'fp32' else 4e-2)
    # eval answers
    interp.eval()
    with torch.no_grad():
        out = interp(pbs, False)
    agree = sum(a == b for a, b in zip(out['answer'], ref32['answer']))
    assert agree == len(questions) if mode == 'fp32' else agree >= len(questions) - 1
