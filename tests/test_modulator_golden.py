"""Attention-transfer calibrator on CPU: program compiler plan + token-side modulator network
(dfol_vqa_b200/modulator.py, plain torch, device-agnostic) + the oracle's apply_modulations restatement, held to
fixtures recorded from the unmodified reference with ``activate_attention_transfer: True``."""

import pytest
import torch

import helpers
import dfol_oracle as orc
from dfol_vqa_b200.compiler import ProgramCompiler
from dfol_vqa_b200.modulator import AttentionTransfer

FILES = helpers.golden_mod_files()


def oracle_with_modulations(case, dtype=torch.float32):
    ont = helpers.ontology_of(case)
    nets = helpers.attention_networks_of(case)
    if dtype == torch.float64:
        nets = [n.double() for n in nets]
    at = AttentionTransfer(nets[0], nets[1], nets[2], ont)
    pbs = helpers.program_batches_of(case, dtype)
    cp = ProgramCompiler(ont, normalize=True, modulated=True).compile(pbs[0], case['counts'])
    rows = at.modulations(cp)
    mods = {(s, k): rows[b:b + r] for s, k, r, b in cp.mod_plan}
    params = {k: v.to(dtype).clone().requires_grad_(True) for k, v in case['state'].items()}
    results, loss = orc.run_step(ont, params, pbs, True, modulations=[mods])
    return results, loss, params, nets, cp


def test_fixtures_present():
    assert len(FILES) >= 8


@pytest.mark.parametrize('path', FILES, ids=lambda p: p.split('goldenmod_')[-1][:-3])
def test_modulated_oracle_matches_reference(path):
    case = helpers.load_golden(path)
    ref = case['ref32']
    results, loss, params, nets, cp = oracle_with_modulations(case)
    lp = results[0]['log_probability']
    assert torch.allclose(lp, ref['log_probability'], rtol=2e-5, atol=2e-6), (lp - ref['log_probability']).abs().max()
    assert abs(float(loss) - float(ref['loss'])) <= 2e-6 * max(1.0, abs(float(ref['loss'])))
    loss.backward()
    for k, p in params.items():
        g = ref['grads'][k]
        assert (p.grad - g).abs().max() <= 2e-5 * g.abs().max() + 1e-7, k
    for net, name in zip(nets, helpers.ATTENTION_NETS):
        for pn, p in net.named_parameters():
            g = ref['grads'][name + '.' + pn]
            mine = p.grad if p.grad is not None else torch.zeros_like(p)
            assert (mine - g).abs().max() <= 2e-5 * g.abs().max() + 2e-7, (name, pn)


@pytest.mark.parametrize('path', FILES[:3], ids=lambda p: p.split('goldenmod_')[-1][:-3])
def test_identity_initialisation(path):
    """With the reference's own initialisation of the output layer (zero weight, bias -> alpha = beta = c = 1, d = .5)
    every modulation is the identity: the modulated run must reproduce the unmodulated log-probabilities."""
    case = helpers.load_golden(path)
    ont = helpers.ontology_of(case)
    from dfol_vqa_b200.networks import build_attention_networks
    nets = build_attention_networks(case['dims']['emb'], case['state_dim'])
    at = AttentionTransfer(nets['forward_attention_network'], nets['backward_attention_network'],
                           nets['attention_output_network'], ont)
    pbs = helpers.program_batches_of(case)
    cp = ProgramCompiler(ont, normalize=True, modulated=True).compile(pbs[0], case['counts'])
    rows = at.modulations(cp)
    assert rows.shape == (cp.mod_rows, 4)
    assert torch.allclose(rows, torch.tensor([0.1, 0.1, 0.1, 0.5]).expand_as(rows), atol=1e-7)
    mods = {(s, k): rows[b:b + r] for s, k, r, b in cp.mod_plan}
    params = {k: v.clone() for k, v in case['state'].items()}
    with torch.no_grad():
        with_mod, _ = orc.run_step(ont, params, pbs, True, modulations=[mods])
        without, _ = orc.run_step(ont, params, pbs, True)
    a, b = with_mod[0]['log_probability'], without[0]['log_probability']
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-5), (a - b).abs().max()
