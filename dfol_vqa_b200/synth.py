"""Synthetic GQA-shaped inputs: question dicts, box features, and the BASELINE.json workloads.

Question dicts follow the reference's program grammar (reference: src/nsvqa/nn/parser/parse_utils.py:196-240,
argument conventions of src/gqa_preprocess.py:292-361, listed in SURVEY.md Appendix A):

  branch    := select [name|'_'] (filter [attr|not(attr)] | relate [rel, is_subject, name|'_'])*
  terminals := exist [] | and [] | or [] | verify_attrs [[a..]] | verify_rel [rel, is_subject, name]
             | choose_attr [[a,b]] | choose_rel [[r1,r2], is_subject, name] | query_attr [category|'name']
             | all_same/all_different/two_same/two_different [category] | compare [attr, is_less]

Object rows follow the box collator's layout (reference: src/nsvqa/data/batch_gqa_boxfeatures_pipeline.py:37-73):
``[feature(D) | img_w, img_h | x, y, w, h]`` with ragged concatenation and an int64 image index per row.
"""

import numpy as np
import torch

TWO_BRANCH = ('and', 'or', 'two_same', 'two_different', 'compare')
BINARY_TERMINALS = ('exist', 'and', 'or', 'verify_attrs', 'verify_rel', 'all_same', 'all_different', 'two_same',
                    'two_different')
ALL_TERMINALS = BINARY_TERMINALS + ('choose_attr', 'choose_rel', 'query_attr', 'compare')


class QuestionSampler(object):

    def __init__(self, ontology, seed=0, neg_prob=0.2, blank_name_prob=0.3):
        self.ont = ontology
        self.rng = np.random.RandomState(seed)
        self.neg_prob = neg_prob
        self.blank_name_prob = blank_name_prob
        self.nouns = [n for n in ontology._nouns if n in ontology._vocabulary['arg_to_idx']]
        self.adjs = [a for a in ontology._adjectives if a in ontology._vocabulary['arg_to_idx']]
        self.rels = [r for r in ontology._relations if r in ontology._vocabulary['arg_to_idx']]
        self.cats = [c for c, v in ontology._attribute_dict.items()
                     if all(m in ontology._vocabulary['arg_to_idx'] for m in v)]

    def _pick(self, seq):
        return seq[self.rng.randint(len(seq))]

    def _name(self):
        return '_' if self.rng.rand() < self.blank_name_prob else self._pick(self.nouns)

    def _attr(self, allow_neg=True):
        a = self._pick(self.adjs if self.rng.rand() < 0.7 else self.nouns)
        return 'not(%s)' % a if allow_neg and self.rng.rand() < self.neg_prob else a

    def _branch(self, hops, relate_prob):
        ops = [{'operator': 'select', 'arguments': [self._name()]}]
        for _ in range(hops):
            if self.rng.rand() < relate_prob:
                ops.append({'operator': 'relate',
                            'arguments': [self._pick(self.rels), bool(self.rng.rand() < 0.5), self._name()]})
            else:
                ops.append({'operator': 'filter', 'arguments': [self._attr()]})
        return ops

    def question(self, terminal, hops, relate_prob=0.35):
        """``hops`` = number of filter/relate ops (a terminal relate counts as one more hop)."""
        rng = self.rng
        nb = 2 if terminal in TWO_BRANCH else 1
        split = [hops // nb + (1 if i < hops % nb else 0) for i in range(nb)]
        branches = [self._branch(h, relate_prob) for h in split]
        answer = 'yes' if rng.rand() < 0.5 else 'no'
        if terminal in ('exist', 'and', 'or'):
            args = []
        elif terminal == 'verify_attrs':
            args = [[self._attr() for _ in range(1 + rng.randint(2))]]
        elif terminal == 'verify_rel':
            args = [self._pick(self.rels), bool(rng.rand() < 0.5), self._name()]
        elif terminal == 'choose_attr':
            cat = self.ont._attribute_dict[self._pick(self.cats)]
            two = [cat[i] for i in rng.choice(len(cat), size=2, replace=False)]
            args = [two]
            answer = two[rng.randint(2)]
        elif terminal == 'choose_rel':
            two = [self.rels[i] for i in rng.choice(len(self.rels), size=2, replace=False)]
            args = [two, bool(rng.rand() < 0.5), self._name()]
            answer = two[rng.randint(2)]
        elif terminal == 'query_attr':
            if rng.rand() < 0.25:
                args = ['name']
                answer = self._pick(self.nouns)
            else:
                cat = self._pick(self.cats)
                args = [cat]
                answer = self._pick(self.ont._attribute_dict[cat])
        elif terminal in ('all_same', 'all_different', 'two_same', 'two_different'):
            args = [self._pick(self.cats)]
        elif terminal == 'compare':
            args = [self._attr(allow_neg=False), bool(rng.rand() < 0.5)]
            n0, n1 = branches[0][0]['arguments'][0], branches[1][0]['arguments'][0]
            answer = (n0 if rng.rand() < 0.5 else n1)
            answer = 'entity' if answer == '_' else answer
        else:
            raise ValueError(terminal)
        tokens = []
        return {'program': {'branches': branches, 'last_op': {'operator': terminal, 'arguments': args}},
                'answer': answer, 'tokens': tokens, 'question': '', 'image_id': '0', 'question_id': '0',
                'original_dict': None}


def make_questions(ontology, batch, terminal, min_hops, max_hops, seed=0, relate_prob=0.35, neg_prob=0.2):
    s = QuestionSampler(ontology, seed=seed, neg_prob=neg_prob)
    return [s.question(terminal, int(s.rng.randint(min_hops, max_hops + 1)), relate_prob) for _ in range(batch)]


def make_relation_chain_questions(ontology, batch, relates, seed=0):
    """select -> relate x ``relates`` -> exist (BASELINE config 3: relation-heavy long programs)."""
    s = QuestionSampler(ontology, seed=seed)
    out = []
    for _ in range(batch):
        ops = [{'operator': 'select', 'arguments': [s._name()]}]
        for _ in range(relates):
            ops.append({'operator': 'relate', 'arguments': [s._pick(s.rels), bool(s.rng.rand() < 0.5), s._name()]})
        out.append({'program': {'branches': [ops], 'last_op': {'operator': 'exist', 'arguments': []}},
                    'answer': 'yes' if s.rng.rand() < 0.5 else 'no', 'tokens': [], 'question': '', 'image_id': '0',
                    'question_id': '0', 'original_dict': None})
    return out


def make_object_features(object_nums, feature_dim, seed=0, feature_scale=1.0, dtype=torch.float32):
    """Ragged box features (T, feature_dim + 6) and the image index of each row (T,)."""
    g = torch.Generator().manual_seed(seed)
    object_nums = [int(n) for n in object_nums]
    total = sum(object_nums)
    feats = torch.randn(total, feature_dim, generator=g) * feature_scale
    img_w = torch.full((total, 1), 640.0)
    img_h = torch.full((total, 1), 480.0)
    x = torch.rand(total, 1, generator=g) * 560.0
    y = torch.rand(total, 1, generator=g) * 400.0
    w = 8.0 + torch.rand(total, 1, generator=g) * (640.0 - x - 8.0)
    h = 8.0 + torch.rand(total, 1, generator=g) * (480.0 - y - 8.0)
    rows = torch.cat([feats, img_w, img_h, x, y, w, h], dim=1).to(dtype)
    batch_index = torch.repeat_interleave(torch.arange(len(object_nums)), torch.tensor(object_nums))
    return rows, batch_index


def object_counts(batch, n_max, ragged, seed=0):
    if not ragged:
        return [n_max] * batch
    rng = np.random.RandomState(seed + 17)
    return rng.randint(max(2, n_max // 2), n_max + 1, size=batch).tolist()
