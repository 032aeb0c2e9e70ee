"""The drop-in boundary exposes the reference's module surface for the path (names and arguments of the methods the
reference's own callers use), and refuses CPU tensors loudly instead of falling back (no GPU needed)."""
import inspect

import pytest
import torch

import helpers


def test_module_surface_matches_reference_names():
    from dfol_vqa_b200 import interpreter as m
    # BatchGQABoxFeaturizer.featurize_scene(device, objects_list, batch_index, meta_data)  batch_gqa_boxfeatures_pipeline.py:199
    assert list(inspect.signature(m.FastBoxFeaturizer.featurize_scene).parameters)[:5] == [
        'self', 'device', 'objects_list', 'batch_index', 'meta_data']
    # ClassifierOracle.compute_all_log_likelihood_2(object_features, pair_object_features)  classifier_oracle.py:145
    assert list(inspect.signature(m.FastClassifierOracle.compute_all_log_likelihood_2).parameters) == [
        'self', 'object_features', 'pair_object_features']
    # OracleBase.get_embedding(tokens, meta_data, device)  base_oracle.py:45
    assert list(inspect.signature(m.FastClassifierOracle.get_embedding).parameters) == [
        'self', 'tokens', 'meta_data', 'device']
    # BatchInterpreterBase.build_scene(device, object_features, batch_index, meta_data) / forward(...)  batch_base_interpreter.py:45,72
    assert list(inspect.signature(m.FastGQAInterpreter.build_scene).parameters)[:5] == [
        'self', 'device', 'object_features', 'batch_index', 'meta_data']
    assert list(inspect.signature(m.FastGQAInterpreter.forward).parameters) == [
        'self', 'program_batch_list', 'is_training', 'return_trace', 'modulator_switch']


def test_inner_surfaces_refuse_cpu_tensors():
    path = [p for p in helpers.golden_files() if 'verify_rel' in p][0]
    case = helpers.load_golden(path)
    ont = helpers.ontology_of(case)
    interp = helpers.build_interpreter(ont, case['dims'], case['state'], device='cpu')
    with pytest.raises(AssertionError):
        interp._featurizer.featurize_scene('cpu', case['features'], case['batch_index'], {})
    with pytest.raises(AssertionError):
        interp._oracle.compute_all_log_likelihood_2(torch.zeros(4, case['dims']['feat'] + 4), None)
    emb = interp._oracle.get_embedding(['red', 'to the left of'], None, 'cpu')
    assert emb.shape == (2, case['dims']['emb']) or emb.shape[0] == 2
