"""Concept ontology for the reasoning path (host side).

Mirror of the reference's ``GQAOntology`` (reference: src/nsvqa/nn/interpreter/batch_gqa_ops.py:25-148).
Only the members the hot path touches are provided, under the reference's attribute names so
either object can be handed to the interpreter:

  _vocabulary['arg_to_idx' | 'idx_to_arg']   1-based concept index              (:52-53)
  _relation_index / _relation_reveresed_index  sorted 0-based relation columns   (:58-62)
  _attribute_index, _noun_index, _nouns, _adjectives, _relations                 (:30-31, :55-60)
  query(name), is_noun(), get_embeddings(names)                                  (:114-148)

Differences, on purpose: ``_nouns`` / ``_adjectives`` / ``_relations`` are *sorted* (the reference uses
``list(set(..))`` whose order depends on PYTHONHASHSEED), and an ontology can be built
from in-memory dicts (synthetic vocabularies for tests / bench) as well as from the reference's JSON files.
"""

import json
import zlib

import numpy as np

UNKNOWN = 'UNKNOWN'


def pseudo_glove(word, dim, scale=0.3):
    """Deterministic stand-in for a GloVe row: N(0, scale^2) seeded by crc32(word)."""
    rng = np.random.RandomState(zlib.crc32(word.encode('utf8')) & 0x7FFFFFFF)
    return (rng.standard_normal(dim) * scale).astype(np.float32)


class Ontology(object):

    def __init__(self, attribute_json_path=None, class_json_path=None, vocab_json_file=None, embedding_file=None,
                 relation_json_path=None, frequency_json_path=None, *, attribute_dict=None, class_dict=None,
                 vocabulary=None, relations=None, embedding_dim=300):
        self._attribute_dict = attribute_dict if attribute_dict is not None else json.load(open(attribute_json_path))
        self._class_dict = class_dict if class_dict is not None else json.load(open(class_json_path))
        self._nouns = sorted(set(sum(self._class_dict.values(), [])))
        self._adjectives = sorted(set(sum(self._attribute_dict.values(), [])))
        self._noun_set = set(self._nouns)

        self._inverted_class_dict = {}
        for parent, members in self._class_dict.items():
            for m in members:
                self._inverted_class_dict.setdefault(m, []).append(parent)

        self._embedding_file = embedding_file
        self._embedding_dim = embedding_dim
        self._word_rows = None  # lazily filled from the embedding file

        if vocabulary is not None:
            self._vocabulary = vocabulary
        else:
            with open(vocab_json_file, 'r') as f:
                self._vocabulary = json.load(f)

        a2i = self._vocabulary['arg_to_idx']
        self._noun_index = sorted(a2i[n] - 1 for n in self._nouns if n in a2i)

        rel = relations
        if rel is None and relation_json_path is not None:
            rel = json.load(open(relation_json_path))
        if rel is not None:
            self._relations = sorted(set(rel))
            self._relation_index = sorted(a2i[r] - 1 for r in self._relations if r in a2i)
            rel_cols = set(self._relation_index)
            self._attribute_index = [i for i in range(len(a2i)) if i not in rel_cols]
            self._attributes = [self._vocabulary['idx_to_arg'][i] for i in self._attribute_index]
            self._relation_reveresed_index = {i: j for j, i in enumerate(self._relation_index)}
            self._attribute_reveresed_index = {i: j for j, i in enumerate(self._attribute_index)}

    # ------------------------------------------------------------------ lookups

    def concept_num(self):
        return len(self._vocabulary['idx_to_arg'])

    def relation_num(self):
        return len(self._relation_index)

    def query(self, name):
        # reference: batch_gqa_ops.py:114-124
        if name in self._attribute_dict:
            return self._attribute_dict[name]
        if name in self._class_dict:
            return self._class_dict[name]
        if name is None:
            return [None]
        if name == 'entity':
            return self._nouns
        return [name]

    def is_noun(self, name):
        return name in self._noun_set

    def is_adjective(self, name):
        return name in self._adjectives

    def is_relation(self, name):
        return name in self._relations

    # --------------------------------------------------------------- embeddings

    def _load_word_rows(self, words):
        want = set(words)
        rows = {}
        with open(self._embedding_file, 'r', encoding='utf8') as f:
            for line in f:
                head = line.split(' ', 1)[0]
                if head in want:
                    rows[head] = np.asarray(line.rstrip().split(' ')[1:], dtype=np.float32)
        return rows

    def get_embeddings(self, names):
        """Sum of per-word vectors of every concept name (reference :135-148).

        With no embedding file the rows are the deterministic pseudo-GloVe vectors above (the real GloVe
        table is not available offline); with a file, rows are read from it (missing words add zero).
        """
        words = [w for n in names for w in n.split(' ')]
        if self._embedding_file is not None:
            rows = self._load_word_rows(words)
            if rows:
                self._embedding_dim = len(next(iter(rows.values())))
        res = np.zeros((len(names), self._embedding_dim), dtype=np.float32)
        for i, name in enumerate(names):
            for w in name.split(' '):
                if self._embedding_file is None:
                    res[i] += pseudo_glove(w, self._embedding_dim)
                elif w in rows:
                    res[i] += rows[w]
        return res


# ---------------------------------------------------------------------- synthetic vocabularies

def synthetic_metadata(concept_num=96, relation_num=12, category_num=4, class_num=3, seed=0):
    """A GQA-shaped vocabulary with generated tokens: returns the four dicts the reference reads from JSON.

    Layout of the 1-based vocabulary: a few control words, attribute-category members ("c<k> v<j>", two-word
    tokens so the embedding sum is exercised), class members (nouns), relations ("rel <k>"), filler concepts.
    Relations are interleaved with the other concepts so that relation columns are not a contiguous range.
    """
    rng = np.random.RandomState(seed)
    budget = concept_num - relation_num
    assert budget >= category_num * 2 + class_num * 2 + 2
    per_cat = max(2, min(8, (budget // 2) // category_num))
    attribute_dict = {'cat%d' % k: ['c%d v%d' % (k, j) for j in range(per_cat - (k % 2))] for k in range(category_num)}
    used = sum(len(v) for v in attribute_dict.values())
    per_cls = max(2, (budget - used - 2) // class_num)
    class_dict = {'class%d' % k: ['noun%d_%d' % (k, j) for j in range(per_cls)] for k in range(class_num)}
    used += sum(len(v) for v in class_dict.values())
    filler = ['misc%d' % i for i in range(budget - used)]
    relations = ['rel %d' % k for k in range(relation_num)]

    tokens = sum(attribute_dict.values(), []) + sum(class_dict.values(), []) + filler
    order = rng.permutation(len(tokens)).tolist()
    tokens = [tokens[i] for i in order]
    # interleave relations at random positions
    pos = sorted(rng.choice(len(tokens) + 1, size=relation_num, replace=False).tolist(), reverse=True)
    for p, r in zip(pos, relations):
        tokens.insert(p, r)
    assert len(tokens) == concept_num and len(set(tokens)) == concept_num
    vocabulary = {'arg_to_idx': {t: i + 1 for i, t in enumerate(tokens)}, 'idx_to_arg': tokens}
    return {'attribute_dict': attribute_dict, 'class_dict': class_dict, 'vocabulary': vocabulary, 'relations': relations}


def synthetic_ontology(concept_num=96, relation_num=12, category_num=4, class_num=3, seed=0, embedding_dim=300):
    md = synthetic_metadata(concept_num, relation_num, category_num, class_num, seed)
    return Ontology(attribute_dict=md['attribute_dict'], class_dict=md['class_dict'], vocabulary=md['vocabulary'],
                    relations=md['relations'], embedding_dim=embedding_dim)
