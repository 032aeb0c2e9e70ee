"""Generates tests/golden/goldenmod_*.pt: the UNMODIFIED reference run with the attention-transfer calibrator ON
(config ``activate_attention_transfer: True``; forward / backward LSTMCells and the output layer randomised so that the
modulations are non-trivial -- the reference initialises the output weight to zero, which makes them constants).

Run here (build container, /root/reference present):  python tests/golden/make_golden_mod.py
Each fixture: synthetic inputs, the reference's oracle state dict and the three attention networks' state dicts, and
what the reference produced in fp32: training-mode log-probabilities, loss/B, gradients of the 12 oracle parameters and
of the 10 attention-network parameters, eval-mode answers.
"""

import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from ref_harness import ReferenceRun, synthetic_metadata  # noqa: E402
from dfol_vqa_b200 import synth  # noqa: E402
from dfol_vqa_b200.ontology import Ontology  # noqa: E402
from make_golden import DIMS, VOCAB  # noqa: E402

STATE_DIM = 8
NETS = ('_forward_attention_network', '_backward_attention_network', '_attention_output_network')
PREFIX = '_ops.select._filter.'


def build_case(terminal, batch, n_max, min_hops, max_hops, seed):
    md = synthetic_metadata(**VOCAB)
    ont = Ontology(attribute_dict=md['attribute_dict'], class_dict=md['class_dict'], vocabulary=md['vocabulary'],
                   relations=md['relations'], embedding_dim=DIMS['emb'])
    questions = synth.make_questions(ont, batch, terminal, min_hops, max_hops, seed=seed)
    counts = synth.object_counts(batch, n_max, True, seed=seed)
    feats, bidx = synth.make_object_features(counts, DIMS['box'], seed=seed + 1)
    run = ReferenceRun(md, DIMS, seed=seed, config_overrides={
        'activate_attention_transfer': True, 'freeze_attention_network': False,
        'attention_transfer_state_dim': STATE_DIM})
    flt = run.model._ops['filter']._filter
    torch.manual_seed(seed + 7)
    with torch.no_grad():
        w = flt._attention_output_network[0].weight
        w.copy_(torch.randn_like(w) * 0.5)
        for net in (flt._forward_attention_network, flt._backward_attention_network):
            for p in net.parameters():
                p.copy_(torch.randn_like(p) * 0.4)
    sd = run.state_dict()
    out = {'terminal': terminal, 'dims': DIMS, 'vocab': VOCAB, 'metadata': md, 'questions': json.dumps(questions),
           'counts': counts, 'features': feats, 'batch_index': bidx, 'split_num': 1, 'state_dim': STATE_DIM,
           'state': {k: v for k, v in sd.items() if k.startswith(('_featurizer.', '_oracle.'))},
           'attention_state': {k[len(PREFIX):]: v for k, v in sd.items()
                               if k.startswith(PREFIX) and k[len(PREFIX):].startswith(NETS)}}
    pbs = run.collate(questions, feats, bidx)
    result, loss, grads = run.loss_and_grads(pbs)
    ev = run.forward(pbs, is_training=False)
    keep = set(out['state']) | {PREFIX + k for k in out['attention_state']}
    rec = {'log_probability': result['log_probability'].detach().clone(), 'loss': loss.clone(),
           'grads': {(k[len(PREFIX):] if k.startswith(PREFIX) else k): g for k, g in grads.items() if k in keep},
           'type': int(result['type']), 'answer': ev['answer'],
           'eval_log_probability': ev['log_probability'].detach().clone()}
    if int(result['type']) == 1:
        rec['options'] = [list(o) for o in result['options']]
    out['ref32'] = rec
    return out


CASES = [
    # terminal, batch, n_max, min_hops, max_hops, seed
    ('exist', 6, 7, 0, 4, 11),
    ('and', 6, 7, 1, 4, 12),
    ('verify_attrs', 6, 7, 0, 3, 14),
    ('verify_rel', 6, 7, 0, 3, 15),
    ('choose_attr', 6, 7, 0, 3, 16),
    ('choose_rel', 6, 7, 0, 3, 17),
    ('query_attr', 5, 6, 0, 3, 18),
    ('all_same', 5, 6, 0, 3, 19),
    ('two_different', 5, 6, 0, 4, 22),
    ('compare', 6, 7, 0, 4, 23),
]


def well_conditioned(case):
    lp = case['ref32']['log_probability']
    return float(lp.max()) < -1e-3 and float(lp.min()) > -12.0


def main():
    for terminal, batch, n_max, lo, hi, seed in CASES:
        for attempt in range(80):
            case = build_case(terminal, batch, n_max, lo, hi, seed + 100 * attempt)
            if well_conditioned(case):
                break
        else:
            raise RuntimeError('no well-conditioned seed for %s' % terminal)
        case['seed'] = seed + 100 * attempt
        name = 'goldenmod_%s.pt' % terminal
        torch.save(case, os.path.join(HERE, name))
        lp = case['ref32']['log_probability']
        print('%-16s lp[%d] range [%.3g, %.3g] loss %.4f  %d bytes' % (
            terminal, lp.numel(), lp.min(), lp.max(), float(case['ref32']['loss']),
            os.path.getsize(os.path.join(HERE, name))))


if __name__ == '__main__':
    main()
