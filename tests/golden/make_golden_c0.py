"""c0-scale fixture recorded from the UNMODIFIED reference: BASELINE.json configs[0] -- B = 32 questions, N = 48 objects,
2048-d box features, 512 / 256 / 300 network widths, the full 2335-concept / 333-relation vocabulary shape, binary
1-3-hop programs (``exist``) -- i.e. the real dimensions instead of the 40-d toy fixtures.

Run here (build container, /root/reference present; needs ~14 GB RAM for the fp64 run):
    python tests/golden/make_golden_c0.py
The fixture stays small because everything large is a pure function of seeds that both sides evaluate identically
(``case_inputs``): the vocabulary, the questions, the box features and the INITIAL WEIGHTS (our factory's init, loaded
into the reference model before the run).  Stored: the reference's log-probabilities, loss, eval answers and -- of each
of the 12 gradient tensors -- norm, max and 4096 entries at seeded positions, in fp32 and fp64.
"""

import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

DIMS = dict(box=2048, feat=512, hidden=256, emb=300)
VOCAB = dict(concept_num=2335, relation_num=333, category_num=31, class_num=53, seed=0)
EMB_BIAS = -4.0   # trained-like operating point (bench.py): keeps the exists-quantifier over 48 objects unsaturated
SAMPLES = 4096


def case_inputs(seed=7, batch=32, n=48, questions=None):
    """Everything both sides regenerate from seeds."""
    from dfol_vqa_b200 import synth
    from dfol_vqa_b200.factory import model_config
    from dfol_vqa_b200.networks import build_networks
    from dfol_vqa_b200.ontology import Ontology, synthetic_metadata
    md = synthetic_metadata(**VOCAB)
    ont = Ontology(attribute_dict=md['attribute_dict'], class_dict=md['class_dict'], vocabulary=md['vocabulary'],
                   relations=md['relations'], embedding_dim=DIMS['emb'])
    feats, bidx = synth.make_object_features([n] * batch, DIMS['box'], seed=seed + 1)
    torch.manual_seed(seed)
    nets = build_networks(model_config(DIMS), ont)
    nets['embedding_network']._network[1].bias.data.fill_(EMB_BIAS)
    names = {'featurizer_network': '_featurizer._featurizer_network', 'attribute_network': '_oracle._attribute_network',
             'relation_network': '_oracle._relation_network', 'embedding_network': '_oracle._embedding_network'}
    state = {}
    for key, prefix in names.items():
        for k, v in nets[key].state_dict().items():
            state[prefix + '.' + k] = v.detach().clone()
    questions = pick_questions(ont, state, feats, bidx, batch, seed) if questions is None else questions
    return md, ont, questions, feats, bidx, state


def pick_questions(ont, state, feats, bidx, batch, seed):
    """The first ``batch`` WELL-CONDITIONED questions of a seeded candidate stream (the same fixed rule as
    make_golden.well_conditioned: -12 < lp < -1e-3 in the fp64 CPU oracle).  The synthetic sampler also emits questions
    that saturate at this operating point (blank select + negated filter: p = 1 exactly; a named relate target nobody
    matches: p clamped at 1e-20): there log(1 - e^x) has no fp32 resolution, the reference's own fp32 and fp64 losses
    differ by 10 %, and a fixture made of them would pin libm rounding, not the algorithm.  Saturated inputs are covered
    by the oracle-vs-CUDA tests with the probability-space tolerance."""
    sys.path.insert(0, os.path.join(REPO, 'oracle'))
    import dfol_oracle as orc
    from dfol_vqa_b200 import synth
    from dfol_vqa_b200.programs import ProgramCollater
    p64 = {k: v.double() for k, v in state.items()}

    def evaluate(qs):
        with torch.no_grad():
            res, _ = orc.run_step(ont, p64, ProgramCollater(1, lambda q: (feats.double(), bidx)).collate(
                json.loads(json.dumps(qs))), is_training=True)
        return res[0]['log_probability']

    # question i is evaluated on image i.  A question's value also depends on the batch it sits in (slot-wide negation
    # round trip, batch_base_ops.py:212-213), so the CHOSEN batch is re-evaluated and its bad slots are re-drawn
    chosen = synth.make_questions(ont, batch, 'exist', 1, 3, seed=seed, relate_prob=0.35)
    for stream in range(1, 40):
        lp = evaluate(chosen)
        bad = [i for i in range(batch) if not (-12.0 < float(lp[i]) < -1e-3)]
        if not bad:
            return chosen
        cand = synth.make_questions(ont, batch, 'exist', 1, 3, seed=seed + 1000 * stream, relate_prob=0.35)
        for i in bad:
            chosen[i] = cand[i]
    raise RuntimeError('no well-conditioned batch found')


def sample_positions(state):
    g = torch.Generator().manual_seed(1234)
    return {k: torch.randint(0, v.numel(), (min(SAMPLES, v.numel()),), generator=g) for k, v in state.items()}


def summarise(grads, pos):
    out = {}
    for k, idx in pos.items():
        g = grads[k].detach().reshape(-1)
        out[k] = {'norm': float(g.double().norm()), 'max': float(g.abs().max()), 'samples': g[idx].clone()}
    return out


def main():
    from ref_harness import ReferenceRun
    md, ont, questions, feats, bidx, state = case_inputs()
    pos = sample_positions(state)
    out = {'dims': DIMS, 'vocab': VOCAB, 'emb_bias': EMB_BIAS, 'questions': json.dumps(questions), 'seed': 7}
    for tag, dtype in (('ref32', torch.float32), ('ref64', torch.float64)):
        run = ReferenceRun(md, DIMS, seed=7, dtype=dtype)
        missing = [k for k in state if k not in run.state_dict()]
        assert not missing, missing
        # (in place through the live state dict: the reference registers the SAME parameters under several module paths,
        # so load_state_dict of a merged dict would copy the stale aliases back over the new values)
        live = run.model.state_dict()
        with torch.no_grad():
            for k, v in state.items():
                live[k].copy_(v.to(dtype))
        pbs = run.collate(questions, feats.to(dtype), bidx, split_num=1)
        result, loss, grads = run.loss_and_grads(pbs)
        ev = run.forward(pbs, is_training=False)
        out[tag] = {'log_probability': result['log_probability'].detach().clone(), 'loss': loss.clone(),
                    'type': int(result['type']), 'answer': ev['answer'],
                    'eval_log_probability': ev['log_probability'].detach().clone(),
                    'grads': summarise({k: g for k, g in grads.items() if k in state}, pos)}
        print(tag, 'loss', float(loss), 'lp range', float(result['log_probability'].min()),
              float(result['log_probability'].max()))
    path = os.path.join(HERE, 'goldenc0_exist.pt')
    torch.save(out, path)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
