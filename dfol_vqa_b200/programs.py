"""Aligned program batches: the host-side input of the interpreter.

These are duck-type mirrors of the reference's ``OperatorBatch`` / ``ProgramBatch`` /
``ProgramCollaterBase`` (reference: src/nsvqa/data/data_pipeline.py:31-290, 626-783) so the parity tests and
the bench can build inputs on a box where the reference is absent.  The interpreter consumes either these
or the reference's own objects: it only reads ``_op_batch_list[i]._op_name / _arguments / _mask /
_is_terminal``, ``_dependencies``, ``_object_features``, ``_object_batch_index``, ``_answers``.

Slot alignment rule (reference ``collate_programs`` :647-746): for every branch, slot 0 is a ``select`` for all
questions; the op of question k that follows r relates and f filters (since the last relate) lands in slot
(r, filter, f) or (r, relate); slots are emitted ordered by r, filters before the relate; one terminal slot per
distinct terminal operator, depending on the last slot of every branch.
"""

import math

import numpy as np
import torch

QUERY_TERMINALS = ('query_attr', 'choose_attr', 'choose_rel')
BINARY, QUERY, STATEMENT = 0, 1, 2


def flatten_options(list_of_lists):
    """[[a,b],None,[c]] -> ([a,b,None,c], [0,0,1,2]) (reference util.flatten_list, util.py:51-56)."""
    flat, owner = [], []
    for q, sub in enumerate(list_of_lists):
        for item in (sub if sub is not None else [None]):
            flat.append(item)
            owner.append(q)
    return flat, owner


class OperatorBatch(object):
    """One aligned op slot: per-position argument columns + a 0/1 participation mask."""

    def __init__(self, op_name, arguments, question_num, is_terminal, mask=None):
        self._op_name = op_name
        self._is_terminal = is_terminal
        self._op_id = None
        self._question_num = question_num

        rows = list(arguments[:question_num]) + [None] * max(0, question_num - len(arguments))
        width = next((len(r) for r in rows if isinstance(r, list)), 0)
        rows = [list(r) if r is not None else [None] * width for r in rows]
        # column-major: _arguments[pos][question]
        self._arguments = [list(col) for col in zip(*rows)] if width > 0 and len(arguments) > 0 else []

        self._predicate_num = question_num
        self._question_index = None
        self._predicate_question_map = None
        if self._arguments and any(isinstance(el, list) and len(el) > 1 for el in self._arguments[0]):
            flat, owner = flatten_options(self._arguments[0])
            self._predicate_num = len(flat)
            if self._predicate_num != question_num:
                self._question_index = torch.tensor(owner, dtype=torch.int64)

        self._mask = None if mask is None else torch.as_tensor(np.asarray(mask), dtype=torch.float32)

    def create_sparse_map(self):
        if self._question_index is not None:
            idx = torch.stack([torch.arange(self._predicate_num, dtype=torch.int64), self._question_index])
            self._predicate_question_map = torch.sparse_coo_tensor(
                idx, torch.ones(self._predicate_num), (self._predicate_num, self._question_num))

    def __repr__(self):
        return 'OperatorBatch(%s, terminal=%s, args=%s)' % (self._op_name, self._is_terminal, self._arguments)


class ProgramBatch(object):

    _serial = 0

    def __init__(self, device, op_batch_list, dependencies, answers, object_features, object_batch_index=None,
                 original_dicts=None, meta_data=None):
        self._op_batch_list = op_batch_list
        self._dependencies = dependencies
        self._answers = answers
        self._object_features = object_features
        self._object_batch_index = None if object_batch_index is None else torch.as_tensor(object_batch_index)
        self._original_dicts = original_dicts
        self._meta_data = meta_data
        self._device = device
        self._batch_size = op_batch_list[0]._question_num if op_batch_list else 0
        ProgramBatch._serial += 1
        for i, ob in enumerate(op_batch_list):
            if ob._op_id is None:
                ob._op_id = '%d:%d' % (ProgramBatch._serial, i)
        last = op_batch_list[-1]._op_name if op_batch_list else None
        self._question_type = QUERY if last in QUERY_TERMINALS else BINARY

    @property
    def device(self):
        return self._device

    def batch_size(self):
        return self._batch_size

    def create_sparse_tensors(self):
        for ob in self._op_batch_list:
            ob.create_sparse_map()

    def pin_memory(self):
        if isinstance(self._object_features, torch.Tensor):
            self._object_features = self._object_features.pin_memory()
        if self._object_batch_index is not None:
            self._object_batch_index = self._object_batch_index.pin_memory()
        if self._staged is not None:
            self._staged = tuple(t.pin_memory() for t in self._staged)
        compiled = getattr(self, '_dfol_compiled', {})
        for cp in compiled.values():  # packed program tables: one async copy per batch
            if not isinstance(cp.blob, torch.Tensor):
                cp.blob = torch.from_numpy(cp.blob).pin_memory()
        if compiled and self._answers is not None:
            if not hasattr(self, '_dfol_targets_host'):
                from .interpreter import targets_of   # loss targets (trainer.py:185-230): collate-time work
                self._dfol_targets_host = torch.from_numpy(targets_of(next(iter(compiled.values())), self._answers))
            self._dfol_targets_host = self._dfol_targets_host.pin_memory()
        return self

    def strip_for_training(self):
        """Drops what FusedTrainStep never reads from a COMPILED batch -- the op-slot objects with their argument
        strings, the answer strings (the loss targets are already a tensor), the per-question option lists -- so that a
        DataLoader worker hands the trainer a few hundred KB of packed tables instead of pickling Python object graphs.
        The result dict of FastGQAInterpreter.forward needs them: do not strip batches meant for evaluation."""
        assert getattr(self, '_dfol_compiled', None), 'strip_for_training: compile the batch first (attach_compiled)'
        self._op_batch_list, self._dependencies, self._original_dicts = [], [], None
        import numpy as np
        for cp in self._dfol_compiled.values():
            cp.options, cp.names, cp.mod_descs, cp.slot_names = [], [], cp.mod_descs if cp.mod_rows else [], []
            # every table the kernels read is in the packed blob: the arrays / tuple lists it was packed from stay behind
            # (the step needs the instruction COUNT only)
            cp.instr = np.empty((cp.instr.shape[0], 0), dtype=np.int32)
            cp.attr_slices = cp.rel_slices = cp.layout_arrays = cp.opts = cp.slot_after = None
        return self

    _staged = None

    def stage_bf16(self, drop_fp32=False):
        """Collate-time option of the tensor-core mode: keep the box features as bf16 (T, D) plus the six fp32 geometry
        columns (T, 6) for the host->device copy.  The tensor-core scene build casts the features to bf16 as its very
        first step (round to nearest even, exactly what ``Tensor.to(torch.bfloat16)`` does here), so the results are
        bit-identical while the copy -- which bounds the end-to-end step at ~55 GB/s of PCIe -- moves half the bytes."""
        f = self._object_features
        D = f.shape[1] - 6
        assert D % 64 == 0, 'bf16 staging needs a feature width that is a multiple of 64'
        self._staged = (f[:, :D].to(torch.bfloat16).contiguous(), f[:, D:].float().contiguous())
        if drop_fp32:
            self._object_features = None
        return self

    def to_cuda(self, device, non_blocking=True):
        bidx = self._object_batch_index.cuda(device, non_blocking=non_blocking)
        if self._staged is not None:
            feats = None
            staged = tuple(t.cuda(device, non_blocking=non_blocking) for t in self._staged)
        else:
            feats = self._object_features.cuda(device, non_blocking=non_blocking)
            staged = None
        pb = ProgramBatch(torch.device('cuda', device) if isinstance(device, int) else device, self._op_batch_list,
                          self._dependencies, self._answers, feats, bidx, self._original_dicts, self._meta_data)
        pb._staged = staged
        pb._batch_size = self._batch_size   # (a batch stripped for training has no op slots to count from)
        # collate-time products of the fused path (compiled bytecode, object counts, targets) travel with the batch
        for key in ('_dfol_compiled', '_dfol_counts', '_dfol_targets'):
            if hasattr(self, key):
                setattr(pb, key, getattr(self, key))
        # the packed program tables travel with the features (same stream, one small copy per compiled variant)
        if hasattr(self, '_dfol_compiled'):
            from .compiler import upload_tables
            dev = pb._device if isinstance(pb._device, torch.device) else torch.device('cuda', device)
            for cp in self._dfol_compiled.values():
                upload_tables(cp, dev)
            answers_t = getattr(self, '_dfol_targets_host', None)
            if answers_t is not None:
                pb._dfol_targets = answers_t.cuda(device, non_blocking=non_blocking)
        pb._dfol_host = self
        # every device tensor this call allocated (on the CURRENT stream, which may be a copy stream): a consumer on
        # another stream must record_stream() all of them (HostStepPipeline does) before the batch can be dropped
        pb._dfol_device_tensors = [t for t in (feats, bidx, getattr(pb, '_dfol_targets', None)) + tuple(staged or ())
                                   if isinstance(t, torch.Tensor)]
        for cp in getattr(self, '_dfol_compiled', {}).values():
            if cp.device_cache is not None:
                pb._dfol_device_tensors.append(cp.device_cache['blob'])
        return pb


def align_programs(questions, starter='select', separator='relate', filler='filter'):
    """Align the programs of a list of question dicts into op slots; returns (slots, dependencies)."""
    n = len(questions)
    branch_num = max(len(q['program']['branches']) for q in questions)
    slots, deps, branch_ends = [], [], []

    for b in range(branch_num):
        first = [q['program']['branches'][b][0] for q in questions]
        slots.append(OperatorBatch(starter, [op['arguments'] if op['operator'] == starter else ['_'] for op in first],
                                   n, False, mask=np.ones(n, dtype=np.float32)))
        deps.append([])

        # key (relates_so_far, is_relate, filter_position) -> per-question argument rows
        table = {}
        for k, q in enumerate(questions):
            rel_seen, fil_seen = 0, 0
            for op in q['program']['branches'][b][1:]:
                if op['operator'] == filler:
                    key = (rel_seen, 0, fil_seen)
                    fil_seen += 1
                elif op['operator'] == separator:
                    key = (rel_seen, 1, 0)
                    rel_seen += 1
                    fil_seen = 0
                else:
                    continue
                entry = table.setdefault(key, ([None] * n, np.zeros(n, dtype=np.float32)))
                entry[0][k] = op['arguments']
                entry[1][k] = 1.0

        for key in sorted(table):
            args, mask = table[key]
            deps.append([len(slots) - 1])
            slots.append(OperatorBatch(separator if key[1] else filler, args, n, False, mask=mask))
        branch_ends.append(len(slots) - 1)

    terminals = {}
    for k, q in enumerate(questions):
        last = q['program']['last_op']
        entry = terminals.setdefault(last['operator'], ([None] * n, np.zeros(n, dtype=np.float32)))
        entry[0][k] = last['arguments']
        entry[1][k] = 1.0
    for name, (args, mask) in terminals.items():
        slots.append(OperatorBatch(name, args, n, True, mask=mask))
        deps.append(list(branch_ends))
    return slots, deps


class ProgramCollater(object):
    """Splits a list of questions into ``split_num`` contiguous program batches (reference :754-783)."""

    def __init__(self, split_num=1, object_source=None, compiler=None):
        self._split_num = split_num
        self._object_source = object_source  # callable(questions) -> (features (T, D+6), batch_index (T,))
        # optional dfol_vqa_b200.compiler.ProgramCompiler: lower the programs to bytecode HERE, i.e. inside the
        # DataLoader worker that collates the batch (5 ms of Python per 256 questions, off the training loop)
        self._compiler = compiler

    def collate(self, questions):
        n = len(questions)
        parts = min(self._split_num, n)
        size = math.ceil(n / parts)
        out = []
        for i in range(parts):
            chunk = questions[i * size:min((i + 1) * size, n)]
            if not chunk:
                break
            slots, deps = align_programs(chunk)
            feats, bidx = self._object_source(chunk) if self._object_source is not None else (None, None)
            pb = ProgramBatch(torch.device('cpu'), slots, deps, [q['answer'] for q in chunk], feats, bidx,
                              [q.get('original_dict') for q in chunk], meta_data=None)
            for ob in pb._op_batch_list:
                ob._op_id = '%d:%s' % (i, ob._op_id)
            if self._compiler is not None:
                attach_compiled(pb, self._compiler)
            out.append(pb)
        return out


def attach_compiled(program_batch, compiler, give_answer=False):
    """Compile ``program_batch`` (ours or the reference's ProgramBatch) and cache the bytecode on it, so that
    FastGQAInterpreter finds it ready; meant to be called from a DataLoader ``collate_fn`` (worker process)."""
    bidx = program_batch._object_batch_index
    counts = torch.bincount(bidx.to(torch.int64)).tolist()
    program_batch._dfol_counts = counts
    cache = getattr(program_batch, '_dfol_compiled', None)
    if cache is None:
        cache = program_batch._dfol_compiled = {}
    key = (bool(give_answer and compiler.hard_mode), compiler.relation_slots, compiler.modulated, compiler.demand_pairs)
    cache[key] = compiler.compile(program_batch, counts, give_answer=give_answer)
    return program_batch
