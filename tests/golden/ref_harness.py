"""Drives the UNMODIFIED reference (imported from /root/reference/src) on synthetic inputs.

Only usable where /root/reference exists (the build container): used by ``make_golden.py`` to produce the
committed fixtures and by ``tests/test_oracle_vs_reference.py`` (skipped elsewhere).  Recipe: SURVEY.md §8(c).
Nothing from the reference is copied; it is imported, run, and its outputs are recorded.
"""

import copy
import json
import logging
import os
import sys
import tempfile
import types

import numpy as np
import torch

REFERENCE_SRC = '/root/reference/src'
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from dfol_vqa_b200.ontology import pseudo_glove, synthetic_metadata  # noqa: E402


def reference_available():
    return os.path.isdir(REFERENCE_SRC)


def _import_reference():
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    if 'h5py' not in sys.modules:
        try:
            import h5py  # noqa: F401
        except ImportError:
            sys.modules['h5py'] = types.ModuleType('h5py')  # only touched for file IO
    import warnings
    warnings.filterwarnings('ignore')
    import gqa_interpreter_experiments as gie
    from nsvqa.data.data_pipeline import ProgramCollaterBase
    from nsvqa.nn.interpreter.batch_gqa_ops import GQAOntology
    from nsvqa.train.trainer import VQATrainer
    return gie, ProgramCollaterBase, GQAOntology, VQATrainer


def default_config(dims):
    cfg = {
        'model_name': 'golden', 'version': 'v0', 'verbose': False, 'model_path': '/tmp/golden_models',
        'box_features_dim': dims['box'], 'oracle_input_dim': dims['feat'], 'oracle_output_dim': 1,
        'word_embedding_dim': dims['emb'], 'classifier_oracle': True, 'featurizer_layers_config': [],
        'attribute_network_layers_config': [dims['hidden']], 'relation_network_layers_config': [dims['hidden']],
        'operator_layers_config': [], 'normalize_oracle': True, 'dropout': 0.0,
        'freeze_featurizer': False, 'freeze_attribute_network': False, 'freeze_relation_network': False,
        'freeze_embedding_network': False, 'activate_attention_transfer': False,
        'attention_transfer_state_dim': 50, 'freeze_attention_network': True, 'trainable_gate': False,
        'likelihood_threshold': 0, 'hard_mode': False, 'gpu_num': 1, 'first_answer': False, 'clip_norm': 0.65,
        'learning_rate': 1e-4, 'weight_decay': 1e-10, 'cpu_cores_num': 8,
    }
    return cfg


class ReferenceRun(object):
    """Builds the reference model on a synthetic vocabulary and runs program batches through it."""

    def __init__(self, metadata, dims, seed=0, dtype=torch.float32, config_overrides=None):
        gie, Collater, GQAOntology, VQATrainer = _import_reference()
        self._tmp = tempfile.mkdtemp(prefix='dfol_md_')
        paths = {}
        for key, name in (('attribute_dict', 'attr.json'), ('class_dict', 'class.json'), ('vocabulary', 'vocab.json'),
                          ('relations', 'rel.json')):
            paths[key] = os.path.join(self._tmp, name)
            with open(paths[key], 'w') as f:
                json.dump(metadata[key], f)

        emb_dim = dims['emb']

        class SynthOntology(GQAOntology):
            def get_embeddings(self, names):
                res = np.zeros((len(names), emb_dim), dtype=np.float32)
                for i, n in enumerate(names):
                    for w in n.split(' '):
                        res[i] += pseudo_glove(w, emb_dim)
                return res

        self.ontology = SynthOntology(paths['attribute_dict'], paths['class_dict'], paths['vocabulary'], None,
                                      relation_json_path=paths['relations'])
        self.config = default_config(dims)
        if config_overrides:
            self.config.update(config_overrides)
        self.logger = logging.getLogger('golden')
        torch.manual_seed(seed)
        exp = gie.GQAObjectBoxExperiment()
        exp._local_rank = 0
        self.model = exp.build_model(self.config, self.ontology, self.logger)
        self.dtype = dtype
        if dtype == torch.float64:
            self.model.double()
        self.trainer = VQATrainer(False, self.config, self.logger, self.ontology)
        self.trainer._model = self.model
        self.trainer._hardset = None
        self._Collater = Collater

    def state_dict(self):
        return {k: v.detach().clone() for k, v in self.model.state_dict().items()}

    def load_state_dict(self, sd):
        self.model.load_state_dict({k: v.to(self.dtype) if v.is_floating_point() else v for k, v in sd.items()})

    def collate(self, questions, features, batch_index, split_num=1):
        n = len(questions)
        counts = torch.bincount(batch_index, minlength=n).tolist()
        starts = np.concatenate([[0], np.cumsum(counts)]).tolist()
        cursor = {'q': 0}
        emb_dim = self.config['word_embedding_dim'] if self.config['activate_attention_transfer'] else 0

        class C(self._Collater):
            def collate_object_features(inner, qs):
                q0 = cursor['q']
                cursor['q'] += len(qs)
                lo, hi = starts[q0], starts[q0 + len(qs)]
                return features[lo:hi].clone(), (batch_index[lo:hi] - q0).clone()

            def collate_meta_data(inner, qs):
                if not emb_dim:
                    return {}
                # attention-transfer runs read world.word_embedding_dim() from the meta data; an empty index sends
                # every token lookup to ontology.get_embeddings (base_oracle.py:45-55)
                return {'index': {}, 'embedding': torch.zeros(1, emb_dim)}

        pbs = C('select', 'relate', 'filter', split_num).collate(copy.deepcopy(questions))
        for pb in pbs:
            pb.create_sparse_tensors()
            if self.dtype == torch.float64:
                pb.to(torch.float64)
        return pbs

    def forward(self, pbs, is_training=True, return_trace=False):
        if is_training:
            self.model.train()
            return self.model(pbs, True, return_trace=return_trace)
        self.model.eval()
        torch.set_default_dtype(self.dtype)
        try:
            with torch.no_grad():
                return self.model(pbs, False, return_trace=return_trace)
        finally:
            torch.set_default_dtype(torch.float32)

    def _forward_train(self, pbs):
        # in fp64 the reference builds some helper tensors (cluster maps) with the default dtype
        torch.set_default_dtype(self.dtype)
        try:
            return self.model(pbs, True)
        finally:
            torch.set_default_dtype(torch.float32)

    def loss_and_grads(self, pbs):
        """One reference training forward+backward (trainer.py:429-436), no optimizer step."""
        self.model.zero_grad()
        self.model.train()
        result = self._forward_train(pbs)
        if self.dtype == torch.float64:
            # VQATrainer._compute_loss hard-codes fp32 targets (trainer.py:193, 227) and cannot run in fp64; the
            # fp64 "truth" run (noise-floor aid only; the fp32 run above is the pin) evaluates the same two
            # formulas (trainer.py:194 and :230) in fp64 here.
            loss = self._loss_fp64(pbs, result)
        else:
            loss = self.trainer._compute_loss(pbs, result)
        total = sum(pb.batch_size() for pb in pbs)
        loss = loss / total
        loss.backward()
        grads = {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p))
                 for k, p in self.model.named_parameters() if p.requires_grad}
        return result, loss.detach(), grads

    def _loss_fp64(self, pbs, result):
        lp = result['log_probability']
        answers = [a for pb in pbs for a in pb._answers]
        if int(result['type']) == 0:
            target = torch.tensor([a in ('yes', 'yeah', 'yep', 'yup', 'aye', 'yea') for a in answers],
                                  dtype=torch.float64)
            return torch.nn.functional.binary_cross_entropy(lp.exp(), target, reduction='sum')
        target = torch.tensor([a == o for a, op in zip(answers, result['options']) for o in op], dtype=torch.float64)
        sizes = [len(op) for op in result['options']]
        parts = torch.split(lp.exp(), sizes)
        return sum(p.sum().clamp(min=1e-20).log() for p in parts) - (target * lp).sum()

    def scene_tables(self, pb):
        """attribute table (T, C) and relation table (P, nR) + pair index triple of one program batch."""
        self.model.eval()
        with torch.no_grad():
            world = self.model.build_scene(pb.device, pb._object_features, pb._object_batch_index, pb._meta_data)
        rel = world._relation_features
        return world._attribute_features.clone(), rel['features'].clone(), [i.clone() for i in rel['index']]
