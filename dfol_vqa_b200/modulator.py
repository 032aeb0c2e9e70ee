"""Attention-transfer calibrator: the token-side recurrent network that produces the per-predicate modulations
(alpha, beta, c, d) the interpreter kernels apply after every filter-like / relate-like sub-operator.

Reference: the "modulator loops" of BatchInterpreterBase.forward (nsvqa/nn/interpreter/batch_base_interpreter.py:87-140),
``transform_attention`` of the operator modules (batch_base_ops.py:407-467, :598-684; batch_gqa_ops.py:185-203, :230-232,
:269-291, :308-310, :337-338, :373-390, :412-413, :475-477, :503-504, :536-537, :610-613, :683-690, :760-764, :782-783),
``_compute_attention_modulations`` (batch_base_ops.py:275-286) and ``BatchAttentionState`` (batch_base_types.py:256-305).

The modulations depend only on the PROGRAM (operator types, predicate word embeddings, masks, subject flags) -- never on
the image -- so this pass works on (questions x state_dim) tensors: a forward LSTMCell chain over the op slots, a
backward LSTMCell chain over the reversed dependencies, and one Linear+Sigmoid per modulated sub-operator.  It is a few
hundred KFLOP per batch of small dense torch ops on the device the networks live on, attached to autograd; the
object-axis work (applying the modulations inside every hop, forward and backward, and reducing their gradients) is in
the CUDA interpreter kernels (program_fwd.cu / program_bwd.cu), which hand ``d loss / d modulations`` back to autograd.

Input: ``CompiledPrograms.mod_descs`` / ``mod_plan`` written by ProgramCompiler(modulated=True).
"""

import numpy as np
import torch

# BatchGQAInterpreter._ops_index (batch_gqa_interpreter.py:66-68); 17 operator classes
OPS_INDEX = {'all_different': 0, 'all_same': 1, 'and': 2, 'choose_attr': 3, 'choose_rel': 4, 'compare': 5, 'end': 6,
             'exist': 7, 'filter': 8, 'or': 9, 'query_attr': 10, 'relate': 11, 'select': 12, 'two_different': 13,
             'two_same': 14, 'verify_attrs': 15, 'verify_rel': 16}
OPS_NUM = 17


def _strip_negation(tok):
    t = tok.strip()
    return t[4:-1] if t.startswith('not(') and t.endswith(')') else t


def _gate(x, y, g):
    """BatchAttentionState.gate (x * g + y * (1 - g) rowwise, batch_base_types.py:283-290) for flags in {0, 1}: x where
    g == 1, y where g == 0 -- one select per tensor instead of four arithmetic launches."""
    keep = (g > 0).unsqueeze(1)
    return (torch.where(keep, x[0], y[0]), torch.where(keep, x[1], y[1]))


class AttentionTransfer(object):

    def __init__(self, forward_network, backward_network, output_network, ontology):
        self.fwd, self.bwd, self.out = forward_network, backward_network, output_network
        self.ont = ontology
        self.state_dim = forward_network.hidden_size
        self._emb = {}

    def parameters(self):
        return [p for net in (self.fwd, self.bwd, self.out) for p in net.parameters()]

    # ---- features: [one-hot operator (17) | 0 = attribute / 1 = relation | word embedding]  (_get_features :265-273)

    def _embedding(self, tok):
        hit = self._emb.get(tok)
        if hit is None:
            hit = self._emb[tok] = np.asarray(self.ont.get_embeddings([tok]), dtype=np.float32)[0]
        return hit

    def _cached(self, key, build):
        """Program-only tensors (feature matrices, row owners, masks, subject flags) are built once per compiled batch
        and device and kept on the CompiledPrograms object (collate-time products, like the bytecode)."""
        hit = self._cache.get(key)
        if hit is None:
            hit = self._cache[key] = build()
        return hit

    def _features(self, op_name, tokens, rel_flag, ref):
        key = ('feat', op_name, rel_flag, id(tokens), str(ref.device), ref.dtype)
        return self._cached(key, lambda: self._build_features(op_name, tokens, rel_flag, ref))

    def _index(self, owner, ref):
        if owner is None:
            return None
        return self._cached(('own', id(owner), str(ref.device)), lambda: torch.as_tensor(owner, device=ref.device))

    def _vector(self, values, ref):
        return self._cached(('vec', id(values), str(ref.device), ref.dtype),
                            lambda: torch.tensor(values, device=ref.device, dtype=ref.dtype))

    def _build_features(self, op_name, tokens, rel_flag, ref):
        f = np.zeros((len(tokens), self.fwd.input_size), dtype=np.float32)
        col = OPS_INDEX[op_name]
        for r, t in enumerate(tokens):
            if t is not None:  # blank predicates keep an all-zero feature row (batch_base_ops.py:442-444)
                f[r, col] = 1.0
                f[r, OPS_NUM] = rel_flag
                f[r, OPS_NUM + 1:] = self._embedding(_strip_negation(t))
        return torch.from_numpy(f).to(device=ref.device, dtype=ref.dtype)

    def _modulation(self, fwd_state, bwd_state):
        return self.out(torch.cat([fwd_state[0], bwd_state[0]], dim=1))

    # ---- sub-operators

    def _filter(self, key, is_forward, state, op_name, tokens, owner, mods, store):
        """FilterBatch.transform_attention (batch_base_ops.py:407-467); ``owner``: question of every predicate row."""
        feats = self._features(op_name, tokens, 0.0, state[0])
        if is_forward:
            old = state if owner is None else (state[0][owner], state[1][owner])          # expand
            new = self.fwd(feats, old)
            store[key] = new
            return new
        mods[key] = self._modulation(store.pop(key), state)
        new = self.bwd(feats, state)
        if owner is not None:                                                                # squeeze
            new = tuple(torch.zeros(self._B, v.shape[1], device=v.device, dtype=v.dtype).index_add_(0, owner, v) for v in new)
        return new

    def _zeros(self, ref_param, rows):
        z = torch.zeros(rows, self.state_dim, device=ref_param.device, dtype=ref_param.dtype)
        return (z, z.clone())

    def _relate_like(self, i, d, is_forward, state, mods, store):
        """GQARelateBatch / GQAChooseRelBatch .transform_attention (batch_gqa_ops.py:373-390, :269-291) on top of
        GQASelectBatch (:185-203) and RelateBatch.transform_attention (batch_base_ops.py:598-684)."""
        ref = self.fwd.weight_ih
        sel, rel = d.get('select'), d.get('relate')
        B = self._B
        if rel is None:
            raise NotImplementedError('relate slot without any relation predicate')
        tokens, owner, _is_subject = rel   # the subject flags cancel out of both passes (see below)
        own = self._index(owner, ref)
        if is_forward:
            x = self._zeros(ref, B)
            if sel is not None:
                x = self._filter((i, 'select'), True, x, d['op'], sel, None, mods, store)
            # subject set = x where is_subject else the incoming state, object set the other way round (:376-377): the
            # LSTM sees their SUM, which is x + state whatever the flag
            agg = (x[0] + state[0], x[1] + state[1])
            if own is not None:
                agg = (agg[0][own], agg[1][own])
            feats = self._features(d['op'], tokens, 1.0, ref)
            new = self.fwd(feats, agg)
            store[(i, 'relate')] = new
            return new  # subject and object states are the same tensor values; the gate of the two is the identity
        # backward pass (:383-390): subject set = the incoming state where is_subject else zero, object set the other way
        # round.  The execution pass keeps the subject posterior where is_subject and the object posterior elsewhere
        # (batch_gqa_ops.py:371 / :254-257), i.e. exactly the role whose backward state is the incoming state: the
        # modulation row that is used is out([forward h | incoming h]) for every flag, and the LSTM sees the sum of the
        # two sets, which is the incoming state.
        fwd_state = store.pop((i, 'relate'))
        mods[(i, 'relate')] = self._modulation(fwd_state, state)
        feats = self._features(d['op'], tokens, 1.0, ref)
        new = self.bwd(feats, state)
        if own is not None:
            new = tuple(torch.zeros(B, v.shape[1], device=v.device, dtype=v.dtype).index_add_(0, own, v) for v in new)
        if sel is not None:
            self._filter((i, 'select'), False, new, d['op'], sel, None, mods, store)
        return new

    def _transform(self, i, d, is_forward, inputs, mods, store):
        op = d['op']
        ref = self.fwd.weight_ih
        if op == 'select':
            if is_forward:
                x = self._zeros(ref, self._B)
                return x if d['select'] is None else self._filter((i, 'select'), True, x, op, d['select'], None, mods,
                                                                  store)
            return inputs[0] if d['select'] is None else self._filter((i, 'select'), False, inputs[0], op, d['select'],
                                                                      None, mods, store)
        if op in ('relate', 'verify_rel', 'choose_rel'):
            return self._relate_like(i, d, is_forward, inputs[0], mods, store)
        if op in ('exist', 'end'):
            return inputs[0]
        if op in ('and', 'or'):
            return (inputs[0], inputs[1])
        fil = d.get('filter')
        if d.get('two'):
            if fil is None:
                return (inputs[0], inputs[1])
            own = self._index(fil[1], ref)
            x1 = self._filter((i, 'filter0'), is_forward, inputs[0], op, fil[0], own, mods, store)
            x2 = self._filter((i, 'filter1'), is_forward, inputs[1], op, fil[0], own, mods, store)
            return (x1, x2)
        if fil is None:
            return inputs[0]
        own = self._index(fil[1], ref)
        return self._filter((i, 'filter'), is_forward, inputs[0], op, fil[0], own, mods, store)

    # ---- the two passes

    def modulations(self, cp):
        """(rows, 4) tensor of raw modulations in ``cp.mod_plan`` row order (autograd-attached)."""
        descs = cp.mod_descs
        n = len(descs)
        ref = self.fwd.weight_ih
        self._B = cp.question_num
        self._cache = cp.mod_cache
        mods, store = {}, {}

        def mask_of(d):
            return None if d['mask'] is None else self._vector(d['mask'], ref)

        trace = []
        for i, d in enumerate(descs):
            inputs = [trace[j] for j in d['deps']]
            x = self._transform(i, d, True, inputs, mods, store)
            if i < n - 1 and inputs and d['mask'] is not None:
                x = _gate(x, inputs[0], mask_of(d))
            trace.append(x)

        consumers = [[] for _ in range(n)]
        for i, d in enumerate(descs):
            for j in d['deps']:
                consumers[j].append(i)
        last = trace[-1]
        if isinstance(last[0], tuple):
            first = tuple((torch.zeros_like(s[0]), torch.zeros_like(s[1])) for s in last)
        else:
            first = ((torch.zeros_like(last[0]), torch.zeros_like(last[1])),)
        back = [None] * n
        for i in range(n - 1, -1, -1):
            d = descs[i]
            if len(consumers[i]) == 1:
                t = back[consumers[i][0]]
                if isinstance(t[0], tuple):
                    inputs = [t[1] if i == n - 2 else t[0]]
                else:
                    inputs = [t]
            else:
                inputs = list(first)
            x = self._transform(i, d, False, inputs, mods, store)
            if d['deps'] and d['mask'] is not None and not isinstance(x[0], tuple) and i != n - 1:
                x = _gate(x, inputs[0], mask_of(d))
            back[i] = x

        parts = []
        for slot_i, key, rows, base in cp.mod_plan:
            m = mods[(slot_i, key)]
            assert m.shape[0] == rows, (slot_i, key, m.shape, rows)
            parts.append(m)
        if not parts:
            return torch.zeros(0, 4, device=ref.device, dtype=ref.dtype)
        return torch.cat(parts, dim=0)
