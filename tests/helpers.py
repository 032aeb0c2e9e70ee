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


def golden_mod_files():
    """Fixtures recorded with the attention-transfer calibrator on (tests/golden/make_golden_mod.py)."""
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, 'goldenmod_*.pt')))


ATTENTION_NETS = ('_forward_attention_network', '_backward_attention_network', '_attention_output_network')


def attention_networks_of(case, device='cpu'):
    """The three attention-transfer networks of a goldenmod fixture (same initial values as the reference run)."""
    from dfol_vqa_b200.networks import build_attention_networks
    nets = build_attention_networks(case['dims']['emb'], case['state_dim'])
    out = []
    for key, name in zip(('forward_attention_network', 'backward_attention_network', 'attention_output_network'),
                         ATTENTION_NETS):
        net = nets[key]
        net.load_state_dict({k[len(name) + 1:]: v for k, v in case['attention_state'].items() if k.startswith(name)})
        out.append(net.to(device))
    return out


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
    err = (x - ref32).abs()
    # saturated entries (p within ~1e-7 of 0 or 1): fp32 log(1 - e^x) has no resolution there and a one-ulp
    # difference in exp/log moves the log-probability by O(1); the answer distribution is what is comparable
    prob_ok = (x.exp() - ref32.exp()).abs() <= 5e-7
    bad = (err > bound) & ~prob_ok
    ratio = torch.where(prob_ok, torch.zeros_like(err), err / bound)
    return not bool(bad.any()), float(ratio.max())


# ---------------------------------------------------------------------------------------- CUDA-side helpers

from dfol_vqa_b200.factory import build_interpreter, model_config  # noqa: E402,F401


def oracle_params(interp, dtype=torch.float32, requires_grad=False):
    """The interpreter's 12 parameters as a CPU dict under the reference's state-dict keys."""
    sd = interp.state_dict()
    keys = [k for k in sd if k.startswith(('_featurizer.', '_oracle.'))]
    return {k: sd[k].detach().cpu().to(dtype).clone().requires_grad_(requires_grad) for k in keys}


def to_cuda(pbs, device=0):
    return [pb.to_cuda(device) for pb in pbs]
