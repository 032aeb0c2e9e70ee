"""CPU oracle for the differentiable-FOL reasoning path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this file; the product package (dfol_vqa_b200) never does and fails loudly without its CUDA library.

A per-image restatement of the reference algorithm in plain torch-CPU ops (fp32 or fp64), differentiable by
autograd.  Parity status: PINNED -- checked against fixtures recorded from the unmodified reference
(tests/golden/*.pt, produced by tests/golden/make_golden.py) by tests/test_oracle_golden.py, and live against
the imported reference by tests/test_oracle_vs_reference.py where /root/reference exists.

The reference evaluates a whole sub-batch with tensors spanning all T objects of the batch; every output only
depends on the question's own image, which is what is restated here.  Reference citations
(file:line under /root/reference/src) are given per function.
"""

import math
import re

import torch
import torch.nn.functional as F

DEFAULT_LL = -30.0
_NEG_RE = re.compile(r"not\((\w|\s)+\)")

P_FEAT_W = '_featurizer._featurizer_network._network.1.weight'
P_FEAT_B = '_featurizer._featurizer_network._network.1.bias'
P_ATTR_W1 = '_oracle._attribute_network._network.1.weight'
P_ATTR_B1 = '_oracle._attribute_network._network.1.bias'
P_ATTR_W2 = '_oracle._attribute_network._network.4.weight'
P_ATTR_B2 = '_oracle._attribute_network._network.4.bias'
P_REL_W1 = '_oracle._relation_network._network.1.weight'
P_REL_B1 = '_oracle._relation_network._network.1.bias'
P_REL_W2 = '_oracle._relation_network._network.4.weight'
P_REL_B2 = '_oracle._relation_network._network.4.bias'
P_EMB_W = '_oracle._embedding_network._network.1.weight'
P_EMB_B = '_oracle._embedding_network._network.1.bias'
PARAM_KEYS = (P_FEAT_W, P_FEAT_B, P_ATTR_W1, P_ATTR_B1, P_ATTR_W2, P_ATTR_B2, P_REL_W1, P_REL_B1, P_REL_W2,
              P_REL_B2, P_EMB_W, P_EMB_B)

BINARY, QUERY, STATEMENT = 0, 1, 2


# ------------------------------------------------------------------ log-space primitives (nsvqa/nn/interpreter/util.py)

def safe_log(x):
    # util.py:22-25
    return x.clamp(min=1e-20).log()


def log_not(x):
    # util.py:35-36
    return safe_log(1.0 - x.exp())


def log_or(a, b):
    # util.py:32-33
    return safe_log(1.0 - (1.0 - a.exp()) * (1.0 - b.exp()))


def log_parametric_not(x, alpha, beta=1):
    # util.py:46-47
    return safe_log(alpha + beta * (1 - 2 * alpha) * x.exp())


def split_negation(token):
    # util.detect_negations, util.py:68-85
    t = token.strip()
    if _NEG_RE.match(t) is not None:
        return True, t[4:-1]
    return False, t


def is_blank(tok):
    # batch_base_ops.py:315
    return tok is None or tok.strip() in ('', '_')


def apply_modulations(log_attention, m):
    """BatchVariableSet.apply_modulations (batch_base_types.py:170-187) for one predicate: ``m`` = the 4 raw outputs
    (alpha, beta, c, d) of the attention output network (gqa_interpreter_experiments.py:119-131; no gate column)."""
    alpha, beta, c, d = m[0] * 10, m[1] * 10, m[2] * 10, m[3]
    temp = alpha * log_attention + safe_log(c) + safe_log(d)
    return temp - safe_log((beta * log_not(log_attention) + safe_log(1.0 - d)).exp() + temp.exp())


# ------------------------------------------------------------------ scene (featurizer + visual oracle)

DROP_FEATURES, DROP_ATTR_IN, DROP_ATTR_HIDDEN, DROP_REL_IN, DROP_REL_HIDDEN, DROP_EMB_ATTR, DROP_EMB_REL = range(7)


def scene_tables(params, features, batch_index, relation_index, masks=None):
    """Per-image attribute and relation log-likelihood tables.

    ``masks`` (training-mode dropout, nn.Dropout in front of every Linear: regular_mlp.py:29-32, embedding_layer.py:73):
    None, or {site: (rows, cols) tensor of 0 / 1/(1-p) scale factors}; pair-level sites hold n_b^2 rows per image in
    (image, subject, object) order, self pairs included.  The reference draws its masks from torch's RNG; parity runs
    feed both sides the masks the CUDA path generates.

    featurize_scene: nsvqa/data/batch_gqa_boxfeatures_pipeline.py:199-281 (featurizer = Linear+Sigmoid,
    gqa_interpreter_experiments.py:28-33 with an empty hidden list); tables: ClassifierOracle.
    compute_all_log_likelihood_2, nsvqa/nn/vision/classifier_oracle.py:145-156.  Dropout is identity (p=0/eval).
    Returns lists over images: attr[b] (N_b, C), rel[b] (N_b, N_b, nR) with rel[b][s, s, :] = -30.
    """
    x = features
    m = masks or {}

    def drop(t, site, rows=None):
        if site not in m:
            return t
        k = m[site].to(t.dtype)
        return t * (k if rows is None else k[rows])

    f = torch.sigmoid(F.linear(drop(x[:, :-6], DROP_FEATURES), params[P_FEAT_W], params[P_FEAT_B]))
    size = torch.stack([x[:, -6], x[:, -5], x[:, -6], x[:, -5]], dim=1).clamp(min=1)
    pos = x[:, -4:] / size
    obj = torch.cat([f, pos], dim=1)

    def head(h, w1, b1, w2, b2, sites, rows=None):
        h = F.elu(F.linear(drop(h, sites[0], rows), params[w1], params[b1]))
        h = torch.sigmoid(F.linear(drop(h, sites[1], rows), params[w2], params[b2]))
        return drop(h, sites[2], rows)

    attr_all = F.logsigmoid(F.linear(head(obj, P_ATTR_W1, P_ATTR_B1, P_ATTR_W2, P_ATTR_B2,
                                          (DROP_ATTR_IN, DROP_ATTR_HIDDEN, DROP_EMB_ATTR)), params[P_EMB_W],
                                     params[P_EMB_B]))
    rel_w = params[P_EMB_W][relation_index]
    rel_b = params[P_EMB_B][relation_index]

    image_num = int(batch_index.max().item()) + 1
    attr, rel = [], []
    pair_start = 0
    for b in range(image_num):
        rows = (batch_index == b).nonzero().flatten()
        n = rows.numel()
        attr.append(attr_all[rows])
        o, p = obj[rows], pos[rows]
        s_idx = torch.arange(n).repeat_interleave(n)
        o_idx = torch.arange(n).repeat(n)
        x1, y1, w1, h1 = (p[s_idx, k] for k in range(4))
        x2, y2, w2, h2 = (p[o_idx, k] for k in range(4))
        dy = y1 + h1 / 2.0 - y2 - h2 / 2.0
        dist = torch.sqrt((x1 + w1 / 2.0 - x2 - w2 / 2.0) ** 2 + dy ** 2)
        # sqrt'(0) = inf on the (unused) diagonal: keep autograd finite by replacing those rows' distance
        off = s_idx != o_idx
        dist = torch.where(off, dist, torch.ones_like(dist))
        ang = torch.asin(dy / dist.clamp(min=1e-10))
        pair = torch.cat([o[s_idx], o[o_idx], dist[:, None], ang[:, None], (x2 - x1).sign()[:, None],
                          (y2 - y1).sign()[:, None]], dim=1)
        pair_rows = slice(pair_start, pair_start + n * n)
        pair_start += n * n
        ll = F.logsigmoid(F.linear(head(pair, P_REL_W1, P_REL_B1, P_REL_W2, P_REL_B2,
                                        (DROP_REL_IN, DROP_REL_HIDDEN, DROP_EMB_REL), pair_rows), rel_w, rel_b))
        ll = torch.where(off[:, None], ll, torch.full_like(ll, DEFAULT_LL))
        rel.append(ll.view(n, n, -1))
    return attr, rel


# ------------------------------------------------------------------ quantifier aggregation

def exists(att):
    """log P(exists) = n(sum_t n(a_t)); BatchVariableSet.log_probability, batch_base_types.py:113-123."""
    return log_not(log_not(att).sum(-1))


def for_all(att):
    # same function with quantifier 0: log_parametric_not(x, 0, 1) = safe_log(exp(x))
    return safe_log(safe_log(att.exp()).sum(-1).exp())


def exists_hard(att):
    # batch_base_types.py:104-112 with quantifier EXISTS
    return log_not(log_not(att).min(-1)[0])


def for_all_hard(att):
    # batch_base_types.py:104-112 with quantifier FOR_ALL: log_parametric_not(x, 0, 1) = safe_log(exp(x)) on both sides
    return safe_log(safe_log(att.exp()).min(-1)[0].exp())


# ------------------------------------------------------------------ the interpreter

class OracleInterpreter(object):
    """Executes aligned program batches the way BatchInterpreterBase.forward does (per question, per image).

    nsvqa/nn/interpreter/batch_base_interpreter.py:72-183, batch_gqa_interpreter.py:72-78, op modules in
    batch_gqa_ops.py / batch_base_ops.py (cited per method).
    """

    def __init__(self, ontology, params, normalize=True, likelihood_threshold=0.0, hard_mode=False):
        self.ont = ontology
        self.params = params
        self.normalize = normalize
        self.threshold = likelihood_threshold
        self.hard_mode = hard_mode
        self.rel_index = torch.tensor(ontology._relation_index, dtype=torch.int64)

    # ---- predicate likelihoods (classifier_oracle.py:44-137 + BatchBayesianLogicCell.forward, batch_base_ops.py:189-213)

    def _attr_col(self, tok):
        return self.ont._vocabulary['arg_to_idx'][tok] - 1

    def _rel_col(self, tok):
        return self.ont._relation_reveresed_index[self.ont._vocabulary['arg_to_idx'][tok] - 1]

    def _predicate_ll(self, table, kind, options, normalise, roundtrip):
        """options: tokens of ONE question (one cluster).  table: attr (N,C) or rel (N,N,nR).  Returns list of
        clamped (and negated) likelihoods, one per option."""
        parsed = [split_negation(t) for t in options]
        cols = [self._attr_col(t) if kind == 'attr' else self._rel_col(t) for _, t in parsed]
        raw = [table[..., c] for c in cols]
        if normalise:
            denom = safe_log(torch.stack([r.exp() for r in raw]).sum(0))
            raw = [r - denom for r in raw]
            if kind == 'rel':  # self pairs are never normalised: they are not in the pair list
                n = table.shape[0]
                eye = torch.eye(n, dtype=torch.bool)
                raw = [torch.where(eye, torch.full_like(r, DEFAULT_LL), r) for r in raw]
        out = []
        for (neg, _), r in zip(parsed, raw):
            ll = -F.relu(-r)
            if roundtrip:
                ll = log_parametric_not(ll, 1.0 if neg else 0.0, 1)
            out.append(ll)
        return out

    @staticmethod
    def _any_negated(token_lists):
        return any(split_negation(t)[0] for toks in token_lists for t in toks if not is_blank(t))

    # ---- ops

    def _filter_slot(self, attr, atts, tokens, mod=None):
        """Plain FilterBatch over one slot (batch_base_ops.py:311-405): one optional token per question.  ``mod``:
        (B, 4) attention-transfer modulations of the slot (:400-402), applied to EVERY row when the slot has at least
        one predicate -- blank rows included."""
        roundtrip = self._any_negated([[t] for t in tokens if t is not None])
        if all(is_blank(t) for t in tokens):
            return list(atts)
        out = []
        for q, (a, t) in enumerate(zip(atts, tokens)):
            x = a if is_blank(t) else a + self._predicate_ll(attr[q], 'attr', [t], False, roundtrip)[0]
            out.append(x if mod is None else apply_modulations(x, mod[q]))
        return out

    def _select_slot(self, attr, names, mod=None):
        # GQASelectBatch, batch_gqa_ops.py:168-183
        blank = [n is None or n.lower() in ('_', 'scene') for n in names]
        zeros = [torch.zeros(a.shape[0], dtype=a.dtype) for a in attr]
        out_names = ['entity' if b else n for b, n in zip(blank, names)]
        if all(blank):
            return zeros, out_names
        return self._filter_slot(attr, zeros, [None if b else n for b, n in zip(blank, names)], mod), out_names

    @staticmethod
    def _relate_core(ll, a_subj, a_obj):
        """Both role posteriors of BatchBayesianLogicCell._forward_core, arity 2, quantifiers EXISTS/EXISTS
        (batch_base_ops.py:62-151).  ll (N,N) indexed [subject, object]."""
        n = ll.shape[0]
        off = 1.0 - torch.eye(n, dtype=ll.dtype)
        inner_s = log_not(ll + a_obj[None, :]) * off      # diagonal zeroed after the inner quantifier (:112)
        out_subj = a_subj + log_not(inner_s.sum(1))
        inner_o = log_not(ll + a_subj[:, None]) * off
        out_obj = a_obj + log_not(inner_o.sum(0))
        return out_subj, out_obj

    def _relate_slot(self, attr, rel, atts, names_in, relations, is_subject, names, mod_sel=None, mod_rel=None):
        # GQARelateBatch, batch_gqa_ops.py:364-371 + RelateBatch.forward, batch_base_ops.py:483-596 (the kept role's
        # posterior carries that role's modulation, :588-594)
        new, new_names = self._select_slot(attr, [n for n in names], mod_sel)
        roundtrip = self._any_negated([[r] for r in relations if r is not None])
        out, out_names = [], []
        for q in range(len(atts)):
            if is_blank(relations[q]):
                out.append(atts[q])  # restored by the interpreter's mask gate (batch_base_interpreter.py:166-167)
                out_names.append(names_in[q])
                continue
            ll = self._predicate_ll(rel[q], 'rel', [relations[q]], False, roundtrip)[0]
            if is_subject[q]:
                res = self._relate_core(ll, new[q], atts[q])[0]
            else:
                res = self._relate_core(ll, atts[q], new[q])[1]
            out.append(res if mod_rel is None else apply_modulations(res, mod_rel[q]))
            out_names.append(new_names[q])
        return out, out_names

    def _options(self, categories, names):
        # batch_gqa_ops.py:305, :583, :655
        return [self.ont.query(c if c not in ('name', 'type') else n) for c, n in zip(categories, names)]

    def _option_filter(self, attr, atts, option_lists, normalized_probability=True, mod=None):
        """FilterBatch with a predicate->question map: per question list of a + ll_k (modulation row = the flattened
        predicate index)."""
        normalise = self.normalize and normalized_probability and any(len(o) > 1 for o in option_lists)
        roundtrip = self._any_negated(option_lists)
        out, row = [], 0
        for q, opts in enumerate(option_lists):
            xs = []
            for ll in self._predicate_ll(attr[q], 'attr', opts, normalise, roundtrip):
                x = atts[q] + ll
                xs.append(x if mod is None else apply_modulations(x, mod[row]))
                row += 1
            out.append(xs)
        return out

    def _agg(self, att, give_answer):
        return exists_hard(att) if (give_answer and self.hard_mode) else exists(att)

    # ---- the program loop

    def run(self, program_batch, is_training=True, tables=None, modulations=None, masks=None):
        """modulations: None, or {(slot index, sub-operator key): (rows, 4) tensor} of attention-transfer modulations
        (keys 'select' / 'filter' / 'relate' / 'filter0' / 'filter1'; rows = questions, or flattened options)."""
        pb = program_batch
        mods = modulations or {}
        feats = pb._object_features
        bidx = pb._object_batch_index.to(torch.int64)
        attr, rel = tables if tables is not None else scene_tables(self.params, feats, bidx, self.rel_index, masks)
        B = len(attr)
        give_answer = not is_training
        trace = []  # per slot: (attentions, names)
        result = None
        slots = pb._op_batch_list
        for i, slot in enumerate(slots):
            deps = pb._dependencies[i]
            inputs = [trace[d] for d in deps]
            args = slot._arguments
            mask = None if slot._mask is None else [float(m) for m in slot._mask]
            name = slot._op_name
            if name == 'select':
                x = self._select_slot(attr, args[0] if args else [None] * B, mods.get((i, 'select')))
            elif name == 'filter':
                x = (self._filter_slot(attr, inputs[0][0], args[0], mods.get((i, 'filter'))), list(inputs[0][1]))
            elif name == 'relate':
                x = self._relate_slot(attr, rel, inputs[0][0], inputs[0][1], args[0], args[1], args[2],
                                      mods.get((i, 'select')), mods.get((i, 'relate')))
            else:
                assert mask is None or all(m > 0 for m in mask), 'one terminal operator per program batch'
                result = self._terminal(name, attr, rel, inputs, args, give_answer, B,
                                        {k[1]: v for k, v in mods.items() if k[0] == i})
                trace.append(None)
                break
            if inputs and mask is not None:  # gate the unaffected questions (batch_base_interpreter.py:166-167)
                x = ([xa if m > 0 else pa for xa, pa, m in zip(x[0], inputs[0][0], mask)],
                     [xn if m > 0 else pn for xn, pn, m in zip(x[1], inputs[0][1], mask)])
            trace.append(x)
        if result is None:  # implicit 'end' (batch_gqa_interpreter.py:75-76, GQAEndBatch :768-780)
            atts, names = trace[-1]
            # (the reference does not forward hard_mode to the implicit end: always the soft quantifier)
            lp = torch.stack([exists(a) for a in atts])
            result = {'log_probability': lp, 'type': STATEMENT, 'options': [],
                      'answer': [[n] for n in names] if give_answer else []}
        result['trace'] = trace
        return result

    def _binary_answer(self, lp, give_answer):
        if not give_answer:
            return []
        return [['yes'] if math.exp(float(v)) > 0.5 else ['no'] for v in lp]

    def _query_answer(self, lp_lists, option_lists, give_answer):
        # util.find_max_ind / unflatten_list, util.py:58-66
        if not give_answer:
            return []
        out = []
        for lps, opts in zip(lp_lists, option_lists):
            p = torch.stack(list(lps)).detach().exp()
            keep = (p == p.max()) & (p > self.threshold)
            out.append([o for o, k in zip(opts, keep.tolist()) if k])
        return out

    def _terminal(self, name, attr, rel, inputs, args, give_answer, B, mods=None):
        # query_attr, all_different and two_different call their inner operator without forwarding hard_mode
        # (batch_gqa_ops.py:306, :628, :703): their quantifiers stay soft under `hard_mode: True`
        hard_ok = name not in ('query_attr', 'all_different', 'two_different')
        agg = lambda a: self._agg(a, give_answer and hard_ok)
        mods = mods or {}
        if name in ('exist', 'end'):
            atts, names = inputs[0]
            lp = torch.stack([agg(a) for a in atts])
            if name == 'end':
                return {'log_probability': lp, 'type': STATEMENT, 'options': [],
                        'answer': [[n] for n in names] if give_answer else []}
            return {'log_probability': lp, 'type': BINARY, 'options': ['no', 'yes'],
                    'answer': self._binary_answer(lp, give_answer)}

        if name in ('and', 'or'):
            # GQAAndBatch / GQAOrBatch, batch_gqa_ops.py:513-567
            v1 = torch.stack([agg(a) for a in inputs[0][0]])
            v2 = torch.stack([agg(a) for a in inputs[1][0]])
            lp = v1 + v2 if name == 'and' else log_or(v1, v2)
            return {'log_probability': lp, 'type': BINARY, 'options': ['no', 'yes'],
                    'answer': self._binary_answer(lp, give_answer)}

        if name == 'verify_attrs':
            # GQAVerifyAttrsBatch, batch_gqa_ops.py:452-473: un-normalised, prior counted once per attribute
            atts, _ = inputs[0]
            per_q = self._option_filter(attr, atts, args[0], normalized_probability=False, mod=mods.get('filter'))
            lp = torch.stack([agg(torch.stack(x).sum(0)) for x in per_q])
            return {'log_probability': lp, 'type': BINARY, 'options': ['no', 'yes'],
                    'answer': self._binary_answer(lp, give_answer)}

        if name == 'verify_rel':
            # GQAVerifyRelBatch, batch_gqa_ops.py:489-501
            atts, names = inputs[0]
            out, _ = self._relate_slot(attr, rel, atts, names, args[0], args[1], args[2], mods.get('select'),
                                       mods.get('relate'))
            lp = torch.stack([agg(a) for a in out])
            return {'log_probability': lp, 'type': BINARY, 'options': ['no', 'yes'],
                    'answer': self._binary_answer(lp, give_answer)}

        if name in ('choose_attr', 'query_attr'):
            # GQAChooseAttrBatch :215-228, GQAQueryAttrBatch :304-306
            atts, names = inputs[0]
            option_lists = args[0] if name == 'choose_attr' else self._options(args[0], names)
            per_q = self._option_filter(attr, atts, option_lists, mod=mods.get('filter'))
            lp_lists = [[agg(x) for x in xs] for xs in per_q]
            lp = torch.stack([v for l in lp_lists for v in l])
            return {'log_probability': lp, 'type': QUERY, 'options': [list(o) for o in option_lists],
                    'answer': self._query_answer(lp_lists, option_lists, give_answer)}

        if name == 'choose_rel':
            # GQAChooseRelBatch, batch_gqa_ops.py:246-267
            atts, names_in = inputs[0]
            option_lists, is_subject, names = args[0], args[1], args[2]
            new, _ = self._select_slot(attr, list(names), mods.get('select'))
            normalise = self.normalize and any(len(o) > 1 for o in option_lists)
            roundtrip = self._any_negated(option_lists)
            lp_lists = []
            mod_rel, prow = mods.get('relate'), 0
            for q, opts in enumerate(option_lists):
                lls = self._predicate_ll(rel[q], 'rel', opts, normalise, roundtrip)
                row = []
                for ll in lls:
                    if is_subject[q]:
                        res = self._relate_core(ll, new[q], atts[q])[0]
                    else:
                        res = self._relate_core(ll, atts[q], new[q])[1]
                    if mod_rel is not None:
                        res = apply_modulations(res, mod_rel[prow])
                    prow += 1
                    row.append(agg(res))
                lp_lists.append(row)
            lp = torch.stack([v for l in lp_lists for v in l])
            return {'log_probability': lp, 'type': QUERY, 'options': [list(o) for o in option_lists],
                    'answer': self._query_answer(lp_lists, option_lists, give_answer)}

        if name in ('all_same', 'all_different'):
            # GQAAllSameBatch :582-608, GQAAllDifferentBatch :627-639
            atts, names = inputs[0]
            option_lists = self._options(args[0], names)
            per_q = self._option_filter(attr, atts, option_lists, mod=mods.get('filter'))
            lps = []
            for q, xs in enumerate(per_q):
                fa = for_all_hard if (give_answer and self.hard_mode and hard_ok) else for_all
                qk = [fa(log_not(atts[q] + log_not(x))) for x in xs]
                lps.append(log_not(log_not(torch.stack(qk)).sum()))
            lp = torch.stack(lps)
            if name == 'all_different':
                lp = log_not(lp)
            return {'log_probability': lp, 'type': BINARY, 'options': ['no', 'yes'],
                    'answer': self._binary_answer(lp, give_answer)}

        if name in ('two_same', 'two_different'):
            # GQATwoSameBatch :654-681, GQATwoDifferentBatch :702-714
            (atts1, names1), (atts2, _) = inputs[0], inputs[1]
            option_lists = self._options(args[0], names1)
            x1 = self._option_filter(attr, atts1, option_lists, mod=mods.get('filter0'))
            x2 = self._option_filter(attr, atts2, option_lists, mod=mods.get('filter1'))
            lps = []
            for q in range(B):
                both = torch.stack([agg(a) + agg(b) for a, b in zip(x1[q], x2[q])])
                lps.append(log_not(log_not(both).sum()))
            lp = torch.stack(lps)
            if name == 'two_different':
                lp = log_not(lp)
            return {'log_probability': lp, 'type': BINARY, 'options': ['no', 'yes'],
                    'answer': self._binary_answer(lp, give_answer)}

        if name == 'compare':
            # GQACompareBatch, batch_gqa_ops.py:730-758
            (atts1, names1), (atts2, names2) = inputs[0], inputs[1]
            x1 = self._filter_slot(attr, atts1, args[0], mods.get('filter0'))
            x2 = self._filter_slot(attr, atts2, args[0], mods.get('filter1'))
            z = torch.stack([torch.stack([agg(a) for a in x1]), torch.stack([agg(a) for a in x2])], dim=1)
            z = F.log_softmax(z, dim=1)
            alpha = torch.tensor([1.0 if f else 0.0 for f in args[1]], dtype=z.dtype)[:, None]
            lp2 = log_parametric_not(z, alpha, 1)
            options = list(zip(names1, names2))
            answer = [[options[q][int(lp2[q].argmax())]] for q in range(B)] if give_answer else []
            return {'log_probability': lp2.reshape(-1), 'type': QUERY, 'options': options, 'answer': answer}

        raise NotImplementedError(name)


# ------------------------------------------------------------------ loss (nsvqa/train/trainer.py:181-262)

YES = ('yes', 'yeah', 'yep', 'yup', 'aye', 'yea')


def compute_loss(results, answer_lists):
    """Summed loss over program batches of one step (caller divides by the question count, trainer.py:434-435).

    ``results``: list of result dicts (one per program batch, same type); ``answer_lists``: their answers.
    """
    kind = results[0]['type']
    lp = torch.cat([r['log_probability'] for r in results])
    if kind == STATEMENT:
        return -lp.sum()
    answers = [a for al in answer_lists for a in al]
    if kind == BINARY:
        target = torch.tensor([1.0 if a in YES else 0.0 for a in answers], dtype=lp.dtype, device=lp.device)
        return F.binary_cross_entropy(lp.exp(), target, reduction='sum')
    options = [o for r in results for o in r['options']]
    target = torch.tensor([1.0 if a == o else 0.0 for a, op in zip(answers, options) for o in op], dtype=lp.dtype,
                          device=lp.device)
    sizes = [len(op) for op in options]
    total = lp.new_zeros(())
    start = 0
    for n in sizes:
        total = total + safe_log(lp[start:start + n].exp().sum())
        start += n
    return total - (target * lp).sum()


def run_step(ontology, params, program_batches, is_training=True, normalize=True, hard_mode=False,
             likelihood_threshold=0.0, modulations=None):
    """Forward over a list of program batches + loss/B (what VQATrainer._train_batch differentiates).
    ``modulations``: optional list (one dict per program batch, see OracleInterpreter.run)."""
    interp = OracleInterpreter(ontology, params, normalize, likelihood_threshold, hard_mode)
    results = [interp.run(pb, is_training, modulations=None if modulations is None else modulations[k])
               for k, pb in enumerate(program_batches)]
    total = sum(pb.batch_size() for pb in program_batches)
    loss = compute_loss(results, [pb._answers for pb in program_batches]) / total
    return results, loss
