"""Live check of the CPU oracle against the UNMODIFIED reference, imported from /root/reference (build container only;
skipped on the GPU box where the reference does not exist). Larger shapes than the committed goldens, fp64."""

import json
import os

import pytest
import torch

import helpers
import dfol_oracle as orc

pytestmark = pytest.mark.skipif(not os.path.isdir('/root/reference/src'), reason='reference not present')


@pytest.mark.parametrize('terminal', ['verify_rel', 'query_attr', 'choose_rel', 'two_different', 'all_same'])
def test_oracle_matches_live_reference_fp64(terminal):
    from ref_harness import ReferenceRun, synthetic_metadata
    from dfol_vqa_b200 import synth
    from dfol_vqa_b200.ontology import Ontology
    from dfol_vqa_b200.programs import ProgramCollater
    dims = dict(box=96, feat=48, hidden=24, emb=32)
    md = synthetic_metadata(220, 30, 6, 5, seed=4)
    ont = Ontology(attribute_dict=md['attribute_dict'], class_dict=md['class_dict'], vocabulary=md['vocabulary'],
                   relations=md['relations'], embedding_dim=dims['emb'])
    questions = synth.make_questions(ont, 5, terminal, 1, 5, seed=9)
    counts = synth.object_counts(5, 11, True, seed=10)
    feats, bidx = synth.make_object_features(counts, dims['box'], seed=11)
    run = ReferenceRun(md, dims, seed=3, dtype=torch.float64)
    pbs = run.collate(questions, feats.double(), bidx)
    result, loss, grads = run.loss_and_grads(pbs)

    state = {k: v for k, v in run.state_dict().items() if k.startswith(('_featurizer.', '_oracle.'))}
    params = {k: v.double().clone().requires_grad_(True) for k, v in state.items()}
    mine = ProgramCollater(1, lambda qs: (feats.double(), bidx)).collate(json.loads(json.dumps(questions)))
    results, my_loss = orc.run_step(ont, params, mine, is_training=True)
    my_loss.backward()
    lp, ref_lp = results[0]['log_probability'].detach(), result['log_probability'].detach()
    if result['type'] == 1 and terminal != 'compare':
        perm, start = [], 0
        for a, b in zip(results[0]['options'], result['options']):
            perm += [start + list(b).index(m) for m in a]
            start += len(b)
        ref_lp = ref_lp[perm]
    assert torch.allclose(lp, ref_lp, rtol=1e-9, atol=1e-11)
    assert abs(float(my_loss) - float(loss)) <= 1e-9 * max(1.0, abs(float(loss)))
    for k, g in grads.items():
        if k in params:
            mine_g = params[k].grad if params[k].grad is not None else torch.zeros_like(params[k])
            assert (mine_g - g).abs().max() <= 1e-9 * g.abs().max().clamp(min=1e-9) + 1e-13, k


def test_drop_in_construction_with_reference_objects():
    """Our modules are constructed from the REFERENCE's own networks / ontology exactly as build_interpreter does
    (gqa_interpreter_experiments.py:200-240): state-dict keys match the reference's, and the program compiler
    consumes the reference's own ProgramBatch objects, producing the same bytecode as from our collater."""
    import numpy as np
    from ref_harness import ReferenceRun, synthetic_metadata
    from dfol_vqa_b200 import synth
    from dfol_vqa_b200.interpreter import FastBoxFeaturizer, FastClassifierOracle, FastGQAInterpreter
    from dfol_vqa_b200.programs import ProgramCollater
    dims = dict(box=40, feat=24, hidden=16, emb=20)
    md = synthetic_metadata(96, 12, 4, 3, seed=0)
    run = ReferenceRun(md, dims, seed=0)
    ref_model = run.model
    ref_oracle = ref_model._oracle
    featurizer = FastBoxFeaturizer(featurizer_network=ref_model._featurizer._featurizer_network)
    oracle = FastClassifierOracle(run.ontology, ref_oracle._attribute_network, ref_oracle._relation_network,
                                  ref_oracle._embedding_network, normalize=True, cached=True)
    interp = FastGQAInterpreter('golden', oracle, run.ontology, featurizer, trainable_gate=False, likelihood_threshold=0,
                                hard_mode=False, attention_transfer_state_dim=50, cached=True)
    mine = {k for k in interp.state_dict() if k.startswith(('_featurizer.', '_oracle.', '_global_step'))}
    theirs = {k for k in ref_model.state_dict() if k.startswith(('_featurizer.', '_oracle.', '_global_step'))}
    assert mine == theirs
    assert interp.parameter_count() == ref_model.parameter_count()

    for terminal in ('verify_rel', 'query_attr', 'two_same', 'choose_rel'):
        questions = synth.make_questions(helpers.ontology_of({'metadata': md, 'dims': dims}), 6, terminal, 0, 4, seed=8)
        counts = synth.object_counts(6, 7, True, seed=2)
        feats, bidx = synth.make_object_features(counts, dims['box'], seed=3)
        ref_pb = run.collate(questions, feats, bidx)[0]
        my_pb = ProgramCollater(1, lambda qs: (feats, bidx)).collate(json.loads(json.dumps(questions)))[0]
        a = interp._compiler.compile(ref_pb, counts)
        b = interp._compiler.compile(my_pb, counts)
        assert np.array_equal(a.instr, b.instr) and np.array_equal(a.q_instr, b.q_instr)
        assert a.kind == b.kind and a.lp_num == b.lp_num
        # 'entity' option order follows the ontology object that is passed in (the reference's is hash-ordered)
        assert [sorted(map(str, o)) for o in a.options] == [sorted(map(str, o)) for o in b.options]


def test_drop_in_construction_with_attention_transfer():
    """sample_config.yaml's arrangement built from the REFERENCE's own modules (activate_attention_transfer: True,
    frozen oracle networks): our interpreter registers the three attention networks under the reference's parameter
    names, every `_ops.*` path we expose exists in the reference's state dict, a checkpoint of one loads into the other,
    and the compiler plan + token-side modulator (torch statement) reproduce the reference's log-probabilities live."""
    from ref_harness import ReferenceRun, synthetic_metadata
    from dfol_vqa_b200 import synth
    from dfol_vqa_b200.interpreter import FastBoxFeaturizer, FastClassifierOracle, FastGQAInterpreter
    from dfol_vqa_b200.modulator import AttentionTransfer
    dims = dict(box=40, feat=24, hidden=16, emb=20)
    md = synthetic_metadata(96, 12, 4, 3, seed=0)
    run = ReferenceRun(md, dims, seed=0, config_overrides={
        'activate_attention_transfer': True, 'freeze_attention_network': False, 'attention_transfer_state_dim': 8,
        'freeze_featurizer': True, 'freeze_attribute_network': True, 'freeze_relation_network': True,
        'freeze_embedding_network': True})
    ref_model = run.model
    flt = ref_model._ops['filter']._filter
    torch.manual_seed(4)
    with torch.no_grad():
        flt._attention_output_network[0].weight.normal_(0.0, 0.5)
    featurizer = FastBoxFeaturizer(featurizer_network=ref_model._featurizer._featurizer_network)
    ref_oracle = ref_model._oracle
    oracle = FastClassifierOracle(run.ontology, ref_oracle._attribute_network, ref_oracle._relation_network,
                                  ref_oracle._embedding_network, normalize=True, cached=True)
    interp = FastGQAInterpreter('golden', oracle, run.ontology, featurizer, trainable_gate=False, likelihood_threshold=0,
                                hard_mode=False, attention_transfer_state_dim=8,
                                forward_attention_network=flt._forward_attention_network,
                                backward_attention_network=flt._backward_attention_network,
                                attention_output_network=flt._attention_output_network,
                                apply_modulation_everywhere=True, cached=True)
    assert interp._has_modulator
    mine, theirs = interp.state_dict(), ref_model.state_dict()
    ours_ops = {k for k in mine if k.startswith('_ops.')}
    assert ours_ops and ours_ops <= set(theirs), sorted(ours_ops - set(theirs))[:5]
    assert {k for k, _ in interp.named_parameters() if 'attention' in k} == \
        {k for k, _ in ref_model.named_parameters() if 'attention' in k}
    assert interp.parameter_count() == ref_model.parameter_count()
    missing, unexpected = interp.load_state_dict(theirs, strict=False)
    assert not missing
    assert all(k.startswith('_ops.') for k in unexpected)   # the reference repeats the oracle under every operator

    questions = synth.make_questions(helpers.ontology_of({'metadata': md, 'dims': dims}), 6, 'verify_rel', 0, 4, seed=8)
    counts = synth.object_counts(6, 7, True, seed=2)
    feats, bidx = synth.make_object_features(counts, dims['box'], seed=3)
    ref_pb = run.collate(questions, feats, bidx)
    ref_lp = run.forward(ref_pb, is_training=True)['log_probability'].detach()
    cp = interp._compiler.compile(ref_pb[0], counts)
    at = AttentionTransfer(flt._forward_attention_network, flt._backward_attention_network,
                           flt._attention_output_network, run.ontology)
    rows = at.modulations(cp)
    mods = {(s, k): rows[b:b + r] for s, k, r, b in cp.mod_plan}
    sd = run.state_dict()
    params = {k: sd[k] for k in orc.PARAM_KEYS}
    with torch.no_grad():
        mine_lp = orc.OracleInterpreter(run.ontology, params).run(ref_pb[0], True, modulations=mods)['log_probability']
    assert torch.allclose(mine_lp, ref_lp, rtol=2e-5, atol=2e-6), (mine_lp - ref_lp).abs().max()


@pytest.mark.parametrize('overrides', [{'hard_mode': True}, {'normalize_oracle': False, 'likelihood_threshold': 0.3}],
                         ids=['hard_mode', 'unnormalised_threshold'])
def test_oracle_eval_options_match_reference_live(overrides):
    """Config switches that the recorded training fixtures do not exercise -- `hard_mode`, `normalize_oracle: False`,
    `likelihood_threshold` -- checked live: eval-mode log-probabilities and answers of every terminal operator."""
    from ref_harness import ReferenceRun, synthetic_metadata
    from dfol_vqa_b200 import synth
    from dfol_vqa_b200.programs import ProgramCollater
    dims = dict(box=40, feat=24, hidden=16, emb=20)
    md = synthetic_metadata(96, 12, 4, 3, seed=0)
    ont = helpers.ontology_of({'metadata': md, 'dims': dims})
    for terminal in ('exist', 'and', 'or', 'verify_attrs', 'verify_rel', 'choose_attr', 'choose_rel', 'query_attr',
                     'all_same', 'all_different', 'two_same', 'two_different', 'compare'):
        questions = synth.make_questions(ont, 6, terminal, 0, 4, seed=21)
        counts = synth.object_counts(6, 7, True, seed=21)
        feats, bidx = synth.make_object_features(counts, dims['box'], seed=22)
        run = ReferenceRun(md, dims, seed=21, config_overrides=overrides)
        ev = run.forward(run.collate(questions, feats, bidx), is_training=False)
        mine = ProgramCollater(1, helpers.slicing_source(feats, bidx)).collate(json.loads(json.dumps(questions)))
        sd = run.state_dict()
        params = {k: sd[k].clone() for k in orc.PARAM_KEYS}
        with torch.no_grad():
            res = orc.OracleInterpreter(ont, params, normalize=overrides.get('normalize_oracle', True),
                                        likelihood_threshold=overrides.get('likelihood_threshold', 0.0),
                                        hard_mode=overrides.get('hard_mode', False)).run(mine[0], is_training=False)
        a, b = res['log_probability'], ev['log_probability']
        if res['type'] == 1 and terminal != 'compare':
            perm, start = [], 0
            for x, y in zip(res['options'], ev['options']):
                perm += [start + list(y).index(m) for m in x]
                start += len(y)
            b = b[perm]
        assert torch.allclose(a, b, rtol=1e-5, atol=2e-6), (terminal, a, b)
        assert [sorted(x) for x in res['answer']] == [sorted(x) for x in ev['answer']], terminal
