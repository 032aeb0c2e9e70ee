"""Construction of the drop-in modules from a model configuration (what GQAObjectBoxExperiment.build_neural_modules +
build_interpreter do in the reference, gqa_interpreter_experiments.py:107-240), for callers without the reference's
experiment file: bench.py, __graft_entry__.smoke(), the tests."""

import torch


def model_config(dims, dropout=0.0):
    return {'box_features_dim': dims['box'], 'oracle_input_dim': dims['feat'], 'word_embedding_dim': dims['emb'],
            'featurizer_layers_config': [], 'attribute_network_layers_config': [dims['hidden']],
            'relation_network_layers_config': [dims['hidden']], 'dropout': dropout}


def build_interpreter(ont, dims, state=None, device='cuda', gemm_mode='fp32', seed=0, emb_bias=None,
                      attention_nets=None, freeze_oracle=False, dropout=0.0, hard_mode=False, normalize=True,
                      likelihood_threshold=0):
    """FastGQAInterpreter over freshly initialised (or fixture) oracle networks."""
    from dfol_vqa_b200.interpreter import FastBoxFeaturizer, FastClassifierOracle, FastGQAInterpreter
    from dfol_vqa_b200.networks import build_networks
    torch.manual_seed(seed)
    nets = build_networks(model_config(dims, dropout), ont)
    if emb_bias is not None:
        # a trained-like operating point: concept probabilities near 0 for most objects, so that the exists-
        # quantifier over ~50 objects does not saturate at p = 1 (random-init logits are ~N(0, 3^2))
        nets['embedding_network']._network[1].bias.data.fill_(emb_bias)
    featurizer = FastBoxFeaturizer(nets['featurizer_network'])
    oracle = FastClassifierOracle(ont, nets['attribute_network'], nets['relation_network'], nets['embedding_network'],
                                  normalize=normalize, cached=True)
    if freeze_oracle:
        for key in ('featurizer_network', 'attribute_network', 'relation_network', 'embedding_network'):
            nets[key].requires_grad_(False)
    fwd, bwd, out = attention_nets if attention_nets is not None else (None, None, None)
    interp = FastGQAInterpreter('model', oracle, ont, featurizer, gemm_mode=gemm_mode, hard_mode=hard_mode,
                                likelihood_threshold=likelihood_threshold,
                                attention_transfer_state_dim=0 if fwd is None else fwd.hidden_size,
                                forward_attention_network=fwd, backward_attention_network=bwd,
                                attention_output_network=out)
    if state is not None:
        missing, unexpected = interp.load_state_dict(state, strict=False)
        assert not unexpected, unexpected
    return interp.to(device)
