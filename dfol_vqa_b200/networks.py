"""Parameter containers of the visual oracle, with the reference's module structure and state-dict keys.

Mirrors RegularMLP / EmbeddingLayer (reference: src/gqa_interpreter_experiments.py:18-77).  They hold parameters
only: the arithmetic runs in libdfol_b200 (their ``forward`` is intentionally absent -- there is no PyTorch
fallback).  The interpreter accepts either these or the reference's own modules (it only reads the Linear layers
of ``_network``).
"""

import torch
import torch.nn as nn


class RegularMLP(nn.Module):
    """[Dropout, Linear, ELU]* + Dropout, Linear, Sigmoid (reference :20-33)."""

    def __init__(self, input_dim, output_dim, layers_config, dropout):
        super(RegularMLP, self).__init__()
        if layers_config is None:
            self._network = None
        else:
            layers, last = [], input_dim
            for width in layers_config:
                layers += [nn.Dropout(dropout), nn.Linear(last, width), nn.ELU()]
                last = width
            layers += [nn.Dropout(dropout), nn.Linear(last, output_dim), nn.Sigmoid()]
            self._network = nn.Sequential(*layers)

    def forward(self, *_):
        raise RuntimeError('dfol_vqa_b200 networks hold parameters only; compute runs in libdfol_b200')


class EmbeddingLayer(nn.Module):
    """Dropout, Linear(embedding_dim -> concepts), LogSigmoid (reference :60-77)."""

    def __init__(self, input_dim, output_dim, dropout, weights=None, biases=None, freeze_bias=False):
        super(EmbeddingLayer, self).__init__()
        linear = nn.Linear(input_dim, output_dim, bias=not freeze_bias)
        if weights is not None:
            linear.weight = nn.Parameter(weights)
        if biases is not None and not freeze_bias:
            linear.bias = nn.Parameter(biases)
        self._network = nn.Sequential(nn.Dropout(dropout), linear, nn.LogSigmoid())

    def forward(self, *_):
        raise RuntimeError('dfol_vqa_b200 networks hold parameters only; compute runs in libdfol_b200')


def linear_layers(module):
    """The nn.Linear layers of a RegularMLP / EmbeddingLayer (ours or the reference's), in order."""
    net = getattr(module, '_network', None)
    if net is None:
        return []
    return [m for m in net if isinstance(m, nn.Linear)]


def dropout_p(module):
    net = getattr(module, '_network', None)
    if net is None:
        return 0.0
    ps = [m.p for m in net if isinstance(m, nn.Dropout)]
    return max(ps) if ps else 0.0


def build_networks(config, ontology):
    """The four oracle networks of GQAObjectBoxExperiment.build_neural_modules (reference :107-182),
    classifier-oracle branch, with the reference's initialisation (GloVe-summed embedding weight, zero bias)."""
    featurizer = RegularMLP(config['box_features_dim'], config['oracle_input_dim'], config['featurizer_layers_config'],
                            config['dropout'])
    attribute = RegularMLP(config['oracle_input_dim'] + 4, config['word_embedding_dim'],
                           config['attribute_network_layers_config'], config['dropout'])
    concept_num = len(ontology._vocabulary['idx_to_arg'])
    emb_in = config['oracle_input_dim'] + 4 if config['attribute_network_layers_config'] is None \
        else config['word_embedding_dim']
    weights = torch.zeros(concept_num, emb_in)
    torch.nn.init.normal_(weights)
    weights[:, :config['word_embedding_dim']] = torch.from_numpy(
        ontology.get_embeddings(ontology._vocabulary['idx_to_arg']))
    embedding = EmbeddingLayer(emb_in, concept_num, config['dropout'], weights, torch.zeros(concept_num),
                               config.get('freeze_embedding_bias', False))
    rel_in = config.get('relation_features_dim', 2 * config['oracle_input_dim'] + 2 * 4 + 4)
    relation = RegularMLP(rel_in, emb_in, config['relation_network_layers_config'], config['dropout'])
    for flag, net in (('freeze_featurizer', featurizer), ('freeze_attribute_network', attribute),
                      ('freeze_relation_network', relation), ('freeze_embedding_network', embedding)):
        if config.get(flag, False):
            net.requires_grad_(False)
    nets = {'featurizer_network': featurizer, 'attribute_network': attribute, 'relation_network': relation,
            'embedding_network': embedding, 'forward_attention_network': None, 'backward_attention_network': None,
            'attention_output_network': None}
    if config.get('activate_attention_transfer', False):
        nets.update(build_attention_networks(config['word_embedding_dim'], config['attention_transfer_state_dim'],
                                             config.get('freeze_attention_network', False)))
    return nets


def build_attention_networks(word_embedding_dim, state_dim, freeze=False):
    """The attention-transfer networks of GQAObjectBoxExperiment.build_neural_modules (reference :112-135): two
    LSTMCells over [17 operator classes | attribute/relation flag | word embedding] and a 4-output Linear+Sigmoid whose
    initial bias makes every modulation the identity (alpha = beta = c = 1, d = 0.5)."""
    import math
    output_dim, max_activation = 4, 10.0
    forward = nn.LSTMCell(word_embedding_dim + 1 + 17, state_dim)
    backward = nn.LSTMCell(word_embedding_dim + 1 + 17, state_dim)
    output = nn.Sequential(nn.Linear(2 * state_dim, output_dim), nn.Sigmoid())
    output[0].weight = nn.Parameter(torch.zeros(output_dim, 2 * state_dim))
    bias = -math.log(max_activation - 1) * torch.ones(output_dim)
    bias[3] = 0
    output[0].bias = nn.Parameter(bias)
    if freeze:
        for net in (forward, backward, output):
            net.requires_grad_(False)
    return {'forward_attention_network': forward, 'backward_attention_network': backward,
            'attention_output_network': output}
