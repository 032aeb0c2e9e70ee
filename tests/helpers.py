"""Shared test helpers: fixture loading and oracle-side evaluation of a golden case."""

import glob
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(REPO, 'tests', 'golden')
sys.path.insert(0, os.path.join(REPO, 'oracle'))

from dfol_vqa_b200.ontology import Ontology  # noqa: E402
from dfol_vqa_b200.programs import ProgramCollater  # noqa: E402


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, 'golden_*.pt')))


def load_golden(path):
    return torch.load(path, weights_only=False)


def ontology_of(case):
    md = case['metadata']
    return Ontology(attribute_dict=md['attribute_dict'], class_dict=md['class_dict'], vocabulary=md['vocabulary'],
                    relations=md['relations'], embedding_dim=case['dims']['emb'])


def slicing_source(features, batch_index):
    """Object source for ProgramCollater: consecutive question chunks own consecutive image blocks."""
    counts = torch.bincount(batch_index).tolist()
    starts = [0]
    for c in counts:
        starts.append(starts[-1] + c)
    cursor = {'q': 0}

    def source(chunk):
        q0 = cursor['q']
        cursor['q'] += len(chunk)
        lo, hi = starts[q0], starts[q0 + len(chunk)]
        return features[lo:hi].clone(), (batch_index[lo:hi] - q0).clone()
    return source


def program_batches_of(case, dtype=torch.float32, questions=None):
    questions = json.loads(case['questions']) if questions is None else questions
    feats = case['features'].to(dtype)
    pbs = ProgramCollater(case['split_num'], slicing_source(feats, case['batch_index'])).collate(questions)
    for pb in pbs:
        pb.create_sparse_tensors()
    return pbs


def close_to_reference(x, ref32, ref64, rtol=1e-5, atol=1e-6, noise_mult=4.0):
    """|x - ref32| <= rtol*|ref| + atol + noise_mult*|ref32 - ref64| elementwise (the last term is the
    reference's own fp32 rounding noise on ill-conditioned log(1-e^x) entries)."""
    x, ref32, ref64 = x.double(), ref32.double(), ref64.double()
    bound = rtol * ref64.abs() + atol + noise_mult * (ref32 - ref64).abs()
    bad = (x - ref32).abs() > bound
    return not bool(bad.any()), float(((x - ref32).abs() / bound).max())
