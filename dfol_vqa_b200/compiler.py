"""Program compiler: lowers an aligned ``ProgramBatch`` (strings, masks, option lists) to the int32 bytecode the
interpreter kernels execute (include/dfol_b200.h, DFOL_OP_* / DFOL_F_*).

What is resolved here, on the host, once per batch (reference behaviour in parentheses):
  * slot participation: ``None`` / '' / '_' predicates and mask-0 questions emit no instruction -- the kernels never
    see them (FilterBatch/RelateBatch pass-through, batch_base_ops.py:315-317, :385; interpreter mask gate,
    batch_base_interpreter.py:166-167);
  * token -> table column (ClassifierOracle, classifier_oracle.py:51-56, :87-96) and ``not(x)`` detection
    (util.detect_negations, util.py:68-85) including the op-slot-wide "round trip" flag: when any predicate of a
    slot is negated the reference passes every other predicate of that slot through slog(exp(.))
    (batch_base_ops.py:212-213);
  * the op-slot-wide option-normalisation flag (ClassifierOracle._build_map, classifier_oracle.py:22-42: the
    per-object softmax over a question's options is applied iff ANY question of the slot has more than one);
  * variable-name tracking, which decides the option sets of query_attr / all_same / two_same
    (batch_gqa_ops.py:171-177, :305, :583, :655);
  * offsets of the compact gradient slices the backward kernel writes, one per (instruction, table operand).
"""

import re

import os

import numpy as np

from . import capi

_NEG_RE = re.compile(r"not\((\w|\s)+\)")

BINARY, QUERY, STATEMENT = 0, 1, 2
K = capi.K  # constants mirrored from include/dfol_b200.h


_SPLIT_CACHE = {}


def _split_neg(token):
    """(negated, bare token) of a predicate token; ``not(x)`` detection as util.detect_negations (util.py:68-85)."""
    hit = _SPLIT_CACHE.get(token)
    if hit is None:
        t = token.strip()
        hit = (True, t[4:-1]) if (t.startswith('not(') and _NEG_RE.match(t) is not None) else (False, t)
        if len(_SPLIT_CACHE) < 100000:
            _SPLIT_CACHE[token] = hit
    return hit


def _blank(tok):
    return tok is None or tok.strip() in ('', '_')


def _roundup4(x):
    return (x + 3) // 4 * 4


class CompiledPrograms(object):
    """Bytecode + result metadata of one program batch."""

    def pack_tables(self):
        _pack_tables(self)

    __slots__ = ('instr', 'q_instr', 'opts', 'lp_num', 'kind', 'options', 'seg', 'names', 'question_num',
                 'g_attr_size', 'g_rel_size', 'attr_slices', 'rel_slices', 'terminal', 'lp_owner', 'device_cache',
                 'alg_bytes', 'slot_wrow', 'img_slot', 'slot_blk', 'rel_slot_size', 'max_slots', 'mod_plan',
                 'mod_descs', 'mod_rows', 'slot_after', 'slot_names', 'mod_cache', 'mod_tok', 'mod_opcol', 'mod_relflag', 'blob',
                 'blob_layout', 'slice_meta', 'layout_arrays', 'layout_meta')


def _slice_table(slices, B):
    """Per-image grouped slice tables of dfol_table_layer_bwd* from (question, column, g offset[, W row]) tuples."""
    if slices:
        arr = np.asarray(slices, dtype=np.int64)
        arr = arr[np.argsort(arr[:, 0], kind='stable')]
        counts = np.bincount(arr[:, 0], minlength=B)
    else:
        arr = np.zeros((0, 4), dtype=np.int64)
        counts = np.zeros(B, dtype=np.int64)
    img_slice = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    pad = arr if arr.shape[0] else np.zeros((1, 4), dtype=np.int64)
    wrow = pad[:, 3] if pad.shape[1] > 3 else pad[:, 1]
    tables = {'goff': pad[:, 2].astype(np.int32), 'col': pad[:, 1].astype(np.int32), 'wrow': wrow.astype(np.int32),
              'img': pad[:, 0].astype(np.int32), 'img_slice': img_slice}
    meta = {'count': int(arr.shape[0]), 'max_per_image': int(counts.max()) if counts.size else 0}
    return tables, meta


def _pack_tables(cp):
    """Everything the kernels read of a compiled batch -- bytecode, option words, result segments, relation-slot tables,
    the per-image slice tables of the table-layer backward -- in ONE byte blob (16-byte aligned sections): one
    host->device copy per batch instead of ~20 small ones.  Built at compile (= collate) time."""
    B = cp.question_num
    arrays = {'instr': cp.instr, 'q_instr': cp.q_instr, 'opts': cp.opts}
    if cp.seg is not None:
        arrays['seg'] = cp.seg
    if cp.img_slot is not None:
        arrays.update(slot_wrow=cp.slot_wrow, img_slot=cp.img_slot, slot_blk=cp.slot_blk)
    for name, a in cp.layout_arrays.items():  # scene layout tables (engine.SceneLayout.of_compiled)
        arrays['lay.' + name] = a
    cp.slice_meta = {'lay': {}}
    for key, slices in (('attr_slices', cp.attr_slices), ('rel_slices', cp.rel_slices)):
        tables, meta = _slice_table(slices, B)
        cp.slice_meta[key] = meta
        for name, a in tables.items():
            arrays[key + '.' + name] = a
    if cp.mod_rows > 0:  # calibrator: per-row token / operator / flag indices, option owners and slot masks
        arrays.update({'mod_tok': cp.mod_tok, 'mod_opcol': cp.mod_opcol, 'mod_relflag': cp.mod_relflag})
        for i, d in enumerate(cp.mod_descs):
            for k in ('filter', 'relate'):
                v = d.get(k)
                if v is not None and v[1] is not None:
                    arrays['mod_owner:%d:%s' % (i, k)] = np.asarray(v[1], dtype=np.int64)
            if d['mask'] is not None and any(m <= 0 for m in d['mask']):
                arrays['mod_mask:%d' % i] = np.asarray(d['mask'], dtype=np.float32)
    layout, off = {}, 0
    for name, a in arrays.items():
        a = np.ascontiguousarray(a)
        arrays[name] = a
        layout[name] = (off, a.nbytes, a.dtype, a.shape)
        off += (a.nbytes + 15) // 16 * 16
    blob = np.zeros(max(off, 16), dtype=np.uint8)
    for name, a in arrays.items():
        o, nb = layout[name][0], layout[name][1]
        blob[o:o + nb] = a.view(np.uint8).reshape(-1)
    cp.blob, cp.blob_layout = blob, layout


def upload_tables(cp, device):
    """Device copies of the tables of a CompiledPrograms (cached on the object): ONE copy of the packed blob (pinned when
    the batch was pinned; issued on the current stream), then typed views."""
    import torch
    if cp.device_cache is None or cp.device_cache['device'] != device:
        host = cp.blob if isinstance(cp.blob, torch.Tensor) else torch.from_numpy(cp.blob)
        blob = host.to(device, non_blocking=True)
        cache = {'device': device, 'blob': blob}
        for name, (off, nbytes, dtype, shape) in cp.blob_layout.items():
            view = blob[off:off + nbytes].view(getattr(torch, np.dtype(dtype).name)).view(*shape)
            if '.' in name:
                key, sub = name.split('.')
                cache.setdefault(key, dict(cp.slice_meta[key]))[sub] = view
            else:
                cache[name] = view
        cp.device_cache = cache
    return cp.device_cache


class ProgramCompiler(object):

    def __init__(self, ontology, normalize=True, hard_mode=False, relation_slots=False, modulated=False,
                 concept_num=None):
        """modulated: attention-transfer modulations are active (FastGQAInterpreter built with the three attention
        networks): every filter-like / relate-like sub-operator of a slot that has at least one non-blank predicate gets
        one row of (alpha, beta, c, d) per predicate (FilterBatch.forward / RelateBatch.forward apply_modulations,
        batch_base_ops.py:400-402, :588-594); the instruction words MOD / MOD2 index those rows.

        relation_slots: demand-driven relation table.  Instead of indexing a dense [nR] relation table, relate /
        choose_rel operands are renumbered per image to *slots* 0..k_b-1 (the distinct relations the image's own
        program uses); the scene build then evaluates only those k_b columns of ClassifierOracle.
        compute_all_log_likelihood_2 (classifier_oracle.py:154 computes all and slices)."""
        self.ont = ontology
        self.normalize = normalize
        self.hard_mode = hard_mode
        self.relation_slots = relation_slots
        # pair rows only for the images whose program reads a relation (needs the slots; tests switch it off to compare)
        self.demand_pairs = os.environ.get('DFOL_DENSE_PAIRS', '0') != '1'
        self.modulated = modulated
        self._rel_concept = list(ontology._relation_index)
        # rows of the embedding layer = columns of the attribute table (the layout tables are packed per batch)
        self.concept_num = concept_num if concept_num is not None else len(ontology._vocabulary['idx_to_arg'])
        self._a2i = ontology._vocabulary['arg_to_idx']
        self._rel_rev = ontology._relation_reveresed_index
        self._option_cache = {}
        self._attr_cache = {}
        self._rel_cache = {}

    def _attr_word(self, tok):
        hit = self._attr_cache.get(tok)
        if hit is None:
            neg, t = _split_neg(tok)
            hit = self._attr_cache[tok] = ((self._a2i[t] - 1), neg)
        return hit

    def _rel_word(self, tok):
        hit = self._rel_cache.get(tok)
        if hit is None:
            neg, t = _split_neg(tok)
            hit = self._rel_cache[tok] = (self._rel_rev[self._a2i[t] - 1], neg)
        return hit

    def _query(self, category, name):
        key = category if category not in ('name', 'type') else name
        return self.ont.query(key)

    def compile(self, program_batch, object_counts, give_answer=False):
        pb = program_batch
        slots = pb._op_batch_list
        deps = pb._dependencies
        B = len(object_counts)
        hard = K.F_HARD if (give_answer and self.hard_mode) else 0
        a_stride = [_roundup4(n) for n in object_counts]
        r_stride = [_roundup4(n * n) for n in object_counts]

        prog = [[] for _ in range(B)]        # per question: list of instruction tuples
        names = [None] * B                   # current variable name per question
        branch_names = []                    # names at the end of each finished branch
        opts = []
        ga = [0]                             # running offsets of the gradient slices
        gr = [0]
        attr_slices = []                     # (question, column, g offset) per attribute operand
        rel_slices = []
        lp_num = 0
        result = {'kind': None, 'options': [], 'seg': None, 'terminal': None, 'lp_owner': None}
        branch_count = 0

        def attr_slice(q, cols):
            """Reserve consecutive gradient slices (stride a_stride[q]) for attribute columns ``cols`` of image q."""
            cols = cols if isinstance(cols, (list, tuple)) else [cols]
            off = ga[0]
            for k, c in enumerate(cols):
                attr_slices.append((q, c, off + k * a_stride[q]))
            ga[0] += len(cols) * a_stride[q]
            return off

        use_slots = self.relation_slots
        slot_of = [dict() for _ in range(B)]  # per image: relation column -> slot

        def rel_operand(q, col):
            """Table column of relation ``col`` inside image q's block (its slot when demand-driven)."""
            if not use_slots:
                return col
            d = slot_of[q]
            if col not in d:
                d[col] = len(d)
            return d[col]

        def rel_slice(q, cols):
            """Reserve gradient slices for relation columns ``cols``; records (image, table column, g offset, W row)."""
            cols = cols if isinstance(cols, (list, tuple)) else [cols]
            off = gr[0]
            for k, c in enumerate(cols):
                wrow = self._rel_concept[c] if use_slots else c
                rel_slices.append((q, rel_operand(q, c), off + k * r_stride[q], wrow))
            gr[0] += len(cols) * r_stride[q]
            return off

        def emit(q, op, flags=0, a0=-1, a1=-1, a2=-1, out=-1, ga0=-1, ga1=-1, grr=-1, mod=-1, mod2=-1, soft=False):
            # soft: query_attr, all_different and two_different call their inner operator WITHOUT forwarding hard_mode
            # (batch_gqa_ops.py:306, :628, :703), so their quantifiers stay soft even under `hard_mode: True`
            h = hard if (op >= K.OP_EXIST and not soft) else 0
            prog[q].append((op, flags | h, a0, a1, a2, out, ga0, ga1, grr, mod, mod2, 0))

        modulated = self.modulated
        mod_plan = []                        # (slot, key, rows, base row), in row order
        mod_descs = []                       # per slot: what the attention-transfer state passes need
        mod_rows = [0]

        def alloc(slot_i, key, rows):
            """Reserve ``rows`` modulation rows for sub-operator ``key`` of slot ``slot_i``; -1 when modulations are off."""
            if not modulated:
                return -1
            base = mod_rows[0]
            mod_plan.append((slot_i, key, rows, base))
            mod_rows[0] += rows
            return base

        def at(base, off):
            return base + off if base >= 0 else -1

        def flat_rows(lists):
            """(flattened tokens, owner question per row, first row of every question) of per-question option lists."""
            flat, owner, start = [], [], []
            for q, l in enumerate(lists):
                start.append(len(flat))
                flat += list(l)
                owner += [q] * len(l)
            return flat, owner, start

        def name_flags(name_tok, roundtrip):
            """(column, flags) of the name select of a relate-like op."""
            if name_tok is None or name_tok.lower() in ('_', 'scene'):
                return -1, 0
            col, neg = self._attr_word(name_tok)
            return col, (K.F_NAME_NEG if neg else 0) | (K.F_NAME_ROUNDTRIP if roundtrip and not neg else 0)

        def option_words(q, option_lists, kind):
            """Append question q's option columns to ``opts``; returns (start, count, columns)."""
            key = (kind, tuple(option_lists))
            hit = self._option_cache.get(key)
            if hit is None:
                words, cols = [], []
                for tok in option_lists:
                    col, neg = self._attr_word(tok) if kind == 'attr' else self._rel_word(tok)
                    words.append(col | (K.OPT_NEG if neg else 0))
                    cols.append(col)
                hit = (words, cols)
                self._option_cache[key] = hit
            start = len(opts)
            if kind == 'rel' and use_slots:
                opts.extend(rel_operand(q, c) | (w & K.OPT_NEG) for w, c in zip(hit[0], hit[1]))
            else:
                opts.extend(hit[0])
            return start, len(hit[0]), hit[1]

        slot_after = []                      # per non-terminal slot: instructions each question has executed after it
        slot_names = []                      # ... and the variable names at that point

        def close_slot():
            slot_after.append([len(p) for p in prog])
            slot_names.append(list(names))

        for i, slot in enumerate(slots):
            if i > 0:
                close_slot()
            name = slot._op_name
            args = slot._arguments
            mask = slot._mask
            if mask is None:
                mask = [1.0] * B
            else:
                mask = mask.tolist() if hasattr(mask, 'tolist') else [float(m) for m in mask]
            terminal = getattr(slot, '_is_terminal', False) or name in capi.TERMINAL_NAMES
            if not terminal and name == 'select':
                if branch_count >= 1:
                    assert branch_count == 1, 'at most two branches per program'
                    branch_names.append(list(names))
                    for q in range(B):
                        emit(q, K.OP_PUSH)
                branch_count += 1
                toks = args[0] if args else [None] * B
                blank = [t is None or t.lower() in ('_', 'scene') for t in toks]
                any_neg = any(_split_neg(t)[0] for t, b in zip(toks, blank) if not b)
                sel_toks = [None if b else t for t, b in zip(toks, blank)]
                base = alloc(i, 'select', B) if not all(blank) else -1
                mod_descs.append({'op': name, 'select': sel_toks if not all(blank) else None})
                for q in range(B):
                    if blank[q]:
                        names[q] = 'entity'
                        emit(q, K.OP_SELECT, 0, -1, mod=at(base, q))
                    else:
                        names[q] = toks[q]
                        col, neg = self._attr_word(toks[q])
                        fl = (K.F_NEG if neg else 0) | (K.F_ROUNDTRIP if any_neg and not neg else 0)
                        emit(q, K.OP_SELECT, fl, col, ga0=attr_slice(q, col), mod=at(base, q))
            elif not terminal and name == 'filter':
                toks = args[0]
                valid = [mask[q] > 0 and not _blank(toks[q]) for q in range(B)]
                # negation detection spans every non-blank predicate of the slot (mask-0 questions carry None)
                any_neg = any(_split_neg(t)[0] for t in toks if not _blank(t))
                fil_toks = [None if _blank(t) else t for t in toks]
                has = any(t is not None for t in fil_toks)
                base = alloc(i, 'filter', B) if has else -1
                mod_descs.append({'op': name, 'filter': (fil_toks, None) if has else None})
                for q in range(B):
                    if valid[q]:
                        col, neg = self._attr_word(toks[q])
                        fl = (K.F_NEG if neg else 0) | (K.F_ROUNDTRIP if any_neg and not neg else 0)
                        emit(q, K.OP_FILTER, fl, col, ga0=attr_slice(q, col), mod=at(base, q))
                    elif mask[q] > 0 and base >= 0:
                        # blank predicate of a participating question in a slot that has predicates: the reference
                        # passes the attention through and still applies the row's modulation (batch_base_ops.py:385-402)
                        emit(q, K.OP_FILTER, 0, -1, mod=at(base, q))
            elif name in ('relate', 'verify_rel'):
                rels, subj, nms = args[0], args[1], args[2]
                any_neg = any(_split_neg(t)[0] for t in rels if not _blank(t))
                name_valid = [not (t is None or t.lower() in ('_', 'scene')) for t in nms]
                any_name_neg = any(_split_neg(t)[0] for t, v in zip(nms, name_valid) if v)
                sel_toks = [t if v else None for t, v in zip(nms, name_valid)]
                rel_toks = [None if _blank(t) else t for t in rels]
                has_sel, has_rel = any(name_valid), any(t is not None for t in rel_toks)
                base_sel = alloc(i, 'select', B) if has_sel else -1
                base_rel = alloc(i, 'relate', B) if has_rel else -1
                mod_descs.append({'op': name, 'select': sel_toks if has_sel else None,
                                  'relate': (rel_toks, None, [1.0 if f else 0.0 for f in subj]) if has_rel else None})
                for q in range(B):
                    if mask[q] > 0 and not _blank(rels[q]):
                        col, neg = self._rel_word(rels[q])
                        ncol, nfl = name_flags(nms[q], any_name_neg)
                        fl = (K.F_NEG if neg else 0) | (K.F_ROUNDTRIP if any_neg and not neg else 0) | nfl
                        fl |= K.F_SUBJECT if subj[q] else 0
                        emit(q, K.OP_RELATE, fl, rel_operand(q, col), ncol,
                             ga1=attr_slice(q, ncol) if ncol >= 0 else -1, grr=rel_slice(q, col),
                             mod=at(base_rel, q), mod2=at(base_sel, q))
                        names[q] = nms[q] if name_valid[q] else 'entity'
                    elif mask[q] > 0:
                        # The reference would turn the attention into select(name) here (RelateBatch passes both roles
                        # through and GQARelateBatch returns the freshly selected one, batch_gqa_ops.py:364-371,
                        # batch_base_ops.py:568-569).  No GQA program has a relate without a relation
                        # (gqa_preprocess.py:292-361): refused loudly rather than evaluated differently.
                        raise NotImplementedError('relate / verify_rel of a participating question without a relation '
                                                  '(blank predicate) is outside the hot path')
                if name == 'verify_rel':
                    assert all(m > 0 for m in mask), 'one terminal operator per program batch'
                    for q in range(B):
                        emit(q, K.OP_EXIST, out=q)
                    lp_num = B
                    result.update(kind=BINARY, terminal=name, lp_owner=list(range(B)))
                    break
            else:
                assert all(m > 0 for m in mask), 'one terminal operator per program batch'
                result['terminal'] = name
                if name in ('exist', 'end', 'and', 'or'):
                    op = {'exist': K.OP_EXIST, 'end': K.OP_EXIST, 'and': K.OP_AND, 'or': K.OP_OR}[name]
                    for q in range(B):
                        emit(q, op, out=q)
                    lp_num = B
                    result.update(kind=STATEMENT if name == 'end' else BINARY, lp_owner=list(range(B)))
                    mod_descs.append({'op': name})
                elif name == 'verify_attrs':
                    lists = args[0]
                    any_neg = any(_split_neg(t)[0] for l in lists for t in l)
                    flat, owner, first = flat_rows(lists)
                    base = alloc(i, 'filter', len(flat))
                    mod_descs.append({'op': name, 'filter': (flat, owner)})
                    for q in range(B):
                        start, cnt, cols = option_words(q, lists[q], 'attr')
                        emit(q, K.OP_VERIFY_ATTRS, K.F_ROUNDTRIP if any_neg else 0, start, cnt, out=q,
                             ga0=attr_slice(q, cols), mod=at(base, first[q]))
                    lp_num = B
                    result.update(kind=BINARY, lp_owner=list(range(B)))
                elif name in ('choose_attr', 'query_attr', 'all_same', 'all_different', 'two_same', 'two_different'):
                    if name == 'choose_attr':
                        lists = args[0]
                    else:
                        base_names = branch_names[0] if name in ('two_same', 'two_different') else names
                        lists = [self._query(c, nm) for c, nm in zip(args[0], base_names)]
                    any_neg = any(_split_neg(t)[0] for l in lists for t in l)
                    norm = self.normalize and any(len(l) > 1 for l in lists)
                    fl = (K.F_ROUNDTRIP if any_neg else 0) | (K.F_NORMALISE if norm else 0)
                    flat, flat_owner, first = flat_rows(lists)
                    two = name in ('two_same', 'two_different')
                    base = alloc(i, 'filter0' if two else 'filter', len(flat))
                    base2 = alloc(i, 'filter1', len(flat)) if two else -1
                    mod_descs.append({'op': name, 'filter': (flat, flat_owner), 'two': two})
                    if name in ('choose_attr', 'query_attr'):
                        seg, owner = [0], []
                        for q in range(B):
                            start, cnt, cols = option_words(q, lists[q], 'attr')
                            emit(q, K.OP_CHOOSE_ATTR, fl, start, cnt, out=lp_num,
                                 ga0=attr_slice(q, cols), mod=at(base, first[q]), soft=(name == 'query_attr'))
                            lp_num += cnt
                            seg.append(lp_num)
                            owner += [q] * cnt
                        result.update(kind=QUERY, options=[list(l) for l in lists], seg=seg, lp_owner=owner)
                    else:
                        op = K.OP_ALL_SAME if name.startswith('all') else K.OP_TWO_SAME
                        if name.endswith('different'):
                            fl |= K.F_NEGATE_RESULT
                        for q in range(B):
                            start, cnt, cols = option_words(q, lists[q], 'attr')
                            emit(q, op, fl, start, cnt, out=q,
                                 ga0=attr_slice(q, cols), mod=at(base, first[q]), mod2=at(base2, first[q]),
                                 soft=name.endswith('different'))
                        lp_num = B
                        result.update(kind=BINARY, lp_owner=list(range(B)))
                elif name == 'choose_rel':
                    lists, subj, nms = args[0], args[1], args[2]
                    any_neg = any(_split_neg(t)[0] for l in lists for t in l)
                    norm = self.normalize and any(len(l) > 1 for l in lists)
                    name_valid = [not (t is None or t.lower() in ('_', 'scene')) for t in nms]
                    any_name_neg = any(_split_neg(t)[0] for t, v in zip(nms, name_valid) if v)
                    seg, owner = [0], []
                    flat, flat_owner, first = flat_rows(lists)
                    sel_toks = [t if v else None for t, v in zip(nms, name_valid)]
                    base_sel = alloc(i, 'select', B) if any(name_valid) else -1
                    base_rel = alloc(i, 'relate', len(flat))
                    mod_descs.append({'op': name, 'select': sel_toks if any(name_valid) else None,
                                      'relate': (flat, flat_owner, [1.0 if f else 0.0 for f in subj])})
                    for q in range(B):
                        start, cnt, cols = option_words(q, lists[q], 'rel')
                        ncol, nfl = name_flags(nms[q], any_name_neg)
                        fl = (K.F_ROUNDTRIP if any_neg else 0) | (K.F_NORMALISE if norm else 0) | nfl
                        fl |= K.F_SUBJECT if subj[q] else 0
                        emit(q, K.OP_CHOOSE_REL, fl, start, cnt, ncol, out=lp_num,
                             ga1=attr_slice(q, ncol) if ncol >= 0 else -1,
                             grr=rel_slice(q, cols), mod=at(base_rel, first[q]), mod2=at(base_sel, q))
                        lp_num += cnt
                        seg.append(lp_num)
                        owner += [q] * cnt
                    result.update(kind=QUERY, options=[list(l) for l in lists], seg=seg, lp_owner=owner)
                elif name == 'compare':
                    toks, less = args[0], args[1]
                    any_neg = any(_split_neg(t)[0] for t in toks if not _blank(t))
                    cmp_toks = [None if _blank(t) else t for t in toks]
                    has = any(t is not None for t in cmp_toks)
                    base = alloc(i, 'filter0', B) if has else -1
                    base2 = alloc(i, 'filter1', B) if has else -1
                    mod_descs.append({'op': name, 'filter': (cmp_toks, None) if has else None, 'two': True})
                    for q in range(B):
                        fl = K.F_IS_LESS if less[q] else 0
                        if _blank(toks[q]):
                            emit(q, K.OP_COMPARE, fl, -1, out=2 * q, mod=at(base, q), mod2=at(base2, q))
                        else:
                            col, neg = self._attr_word(toks[q])
                            fl |= (K.F_NEG if neg else 0) | (K.F_ROUNDTRIP if any_neg and not neg else 0)
                            emit(q, K.OP_COMPARE, fl, col, out=2 * q, ga0=attr_slice(q, col), mod=at(base, q),
                                 mod2=at(base2, q))
                    lp_num = 2 * B
                    result.update(kind=QUERY, options=list(zip(branch_names[0], names)),
                                  seg=list(range(0, 2 * B + 1, 2)), lp_owner=[q for q in range(B) for _ in (0, 1)])
                else:
                    raise NotImplementedError('operator %r is outside the hot path' % name)
                break

        if result['kind'] is None:
            # last slot is not terminal: implicit 'end' (batch_gqa_interpreter.py:75-76).  The reference calls
            # self._ops['end'](op_id, world, x, not is_training, pqm) WITHOUT forwarding hard_mode (GQAEndBatch.forward
            # defaults it to False, batch_gqa_ops.py:768-773): the implicit end is always the soft quantifier
            close_slot()
            for q in range(B):
                emit(q, K.OP_EXIST, out=q, soft=True)
            lp_num = B
            result.update(kind=STATEMENT, terminal='end', lp_owner=list(range(B)))

        for i, d in enumerate(mod_descs):  # state-pass wiring of the slots that were compiled
            d['deps'] = list(deps[i])
            m = slots[i]._mask
            d['mask'] = None if m is None else [float(v) for v in (m.tolist() if hasattr(m, 'tolist') else m)]

        cp = CompiledPrograms()
        q_instr = np.zeros(B + 1, dtype=np.int32)
        q_instr[1:] = np.cumsum([len(p) for p in prog])
        flat = [ins for p in prog for ins in p]
        cp.instr = np.asarray(flat, dtype=np.int32).reshape(-1, K.INSTR_WORDS)
        cp.q_instr = q_instr
        cp.opts = np.asarray(opts if opts else [0], dtype=np.int32)
        cp.lp_num = lp_num
        cp.kind = result['kind']
        cp.options = result['options']
        cp.seg = None if result['seg'] is None else np.asarray(result['seg'], dtype=np.int32)
        cp.names = list(names)
        cp.question_num = B
        cp.g_attr_size = max(ga[0], 1)
        cp.g_rel_size = max(gr[0], 1)
        cp.attr_slices = attr_slices
        cp.rel_slices = rel_slices
        cp.terminal = result['terminal']
        cp.lp_owner = result['lp_owner']
        cp.device_cache = None
        cp.slot_after = np.asarray(slot_after, dtype=np.int64).reshape(len(slot_after), B)
        cp.slot_names = slot_names
        cp.mod_cache = {}
        # per modulation row: vocabulary index of its predicate token (-1 = blank), operator class and attribute /
        # relation flag -- what the calibrator's feature rows are made of (BatchOperatorBase._get_features,
        # batch_base_ops.py:265-273); built here so that it is collate-time work like the bytecode
        from .modulator import OPS_INDEX
        tok = np.full(max(mod_rows[0], 1), -1, dtype=np.int64)
        opcol = np.zeros(max(mod_rows[0], 1), dtype=np.int64)
        relflag = np.zeros(max(mod_rows[0], 1), dtype=np.float32)
        for slot_i, key, rows, base in mod_plan:
            d = mod_descs[slot_i]
            tokens = d['select'] if key == 'select' else (d['relate'][0] if key == 'relate' else d['filter'][0])
            opcol[base:base + rows] = OPS_INDEX[d['op']]
            relflag[base:base + rows] = 1.0 if key == 'relate' else 0.0
            ids = tok[base:base + rows]
            for r, t in enumerate(tokens):
                if t is not None:
                    ids[r] = self._attr_word(t)[0]   # (vocabulary index of the bare token; relations included)
        cp.mod_tok, cp.mod_opcol, cp.mod_relflag = tok, opcol, relflag
        cp.mod_plan = mod_plan
        cp.mod_descs = mod_descs
        cp.mod_rows = mod_rows[0]
        # algorithmic bytes of one forward pass: every table slice read once + the log-probabilities written
        cp.alg_bytes = 4.0 * (sum(object_counts[s[0]] for s in attr_slices) +
                              sum(object_counts[s[0]] ** 2 for s in rel_slices) + lp_num)
        if use_slots:
            n_slots = np.asarray([len(d) for d in slot_of], dtype=np.int64)
            cp.img_slot = np.concatenate([[0], np.cumsum(n_slots)]).astype(np.int32)
            wrow = []
            for d in slot_of:
                wrow += [self._rel_concept[c] for c, _ in sorted(d.items(), key=lambda kv: kv[1])]
            cp.slot_wrow = np.asarray(wrow if wrow else [0], dtype=np.int32)
            blk = np.concatenate([[0], np.cumsum(n_slots * np.asarray(r_stride, dtype=np.int64))])
            cp.slot_blk = blk[:-1].astype(np.int64)
            cp.rel_slot_size = int(max(blk[-1], 1))
            cp.max_slots = int(n_slots.max()) if B else 0
        else:
            cp.img_slot = cp.slot_wrow = cp.slot_blk = None
            cp.rel_slot_size = cp.max_slots = 0
        # scene layout tables, packed with the bytecode.  With relation slots the pair rows are demand-driven too: an
        # image whose program reads no relation likelihood gets none (engine.layout_tables); the dense pair tables ride
        # along for the paths that need every pair row (training-mode dropout)
        from .engine import layout_tables
        C, nR = self.concept_num, len(self._rel_concept)
        mask = (np.diff(cp.img_slot) > 0) if (use_slots and self.demand_pairs) else None
        cp.layout_arrays, cp.layout_meta = layout_tables(object_counts, C, nR, mask)
        if mask is not None:
            dense, dmeta = layout_tables(object_counts, C, nR, None)
            cp.layout_arrays['img_nn_d'], cp.layout_arrays['pair_row_d'] = dense['img_nn'], dense['pair_row']
            cp.layout_arrays['pair_tile_d'] = dense['pair_tile']
            cp.layout_meta['P_dense'], cp.layout_meta['pair_tiles_dense'] = dmeta['P'], dmeta['pair_tiles']
        cp.pack_tables()
        return cp
