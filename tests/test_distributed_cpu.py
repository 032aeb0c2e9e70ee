"""World-size-2 gloo test of the data-parallel host logic (question sharding + flat-bucket gradient all-reduce).

Each rank computes, with the CPU oracle, the gradients of (its shard's summed loss) / (GLOBAL question count) --
exactly what FusedTrainStep does with the CUDA kernels -- re-homes them in a FlatBucket and all-reduces; the result
must equal the single-process gradients of the whole batch (reference semantics: trainer.py:434-435)."""

import json
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers
import dfol_oracle as orc
from dfol_vqa_b200 import synth
from dfol_vqa_b200.ontology import synthetic_ontology
from dfol_vqa_b200.parallel import FlatBucket, shard_questions, shard_range
from dfol_vqa_b200.programs import ProgramCollater

DIMS = dict(box=24, feat=16, hidden=8, emb=12)


def _setup():
    ont = synthetic_ontology(64, 8, 3, 3, seed=2, embedding_dim=DIMS['emb'])
    questions = synth.make_questions(ont, 7, 'verify_rel', 1, 3, seed=4)
    counts = synth.object_counts(7, 6, True, seed=5)
    feats, bidx = synth.make_object_features(counts, DIMS['box'], seed=6)
    from dfol_vqa_b200.networks import build_networks
    torch.manual_seed(1)
    nets = build_networks(helpers.model_config(DIMS), ont)
    names = {'featurizer_network': '_featurizer._featurizer_network', 'attribute_network': '_oracle._attribute_network',
             'relation_network': '_oracle._relation_network', 'embedding_network': '_oracle._embedding_network'}
    params = {}
    for key, prefix in names.items():
        for k, v in nets[key].state_dict().items():
            params[prefix + '.' + k] = v.clone()
    return ont, questions, counts, feats, bidx, params


def _grads(ont, params, questions, counts, feats, bidx, lo, hi, total):
    starts = [0]
    for c in counts:
        starts.append(starts[-1] + c)
    f = feats[starts[lo]:starts[hi]]
    b = bidx[starts[lo]:starts[hi]] - lo
    p = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    pbs = ProgramCollater(1, lambda qs: (f, b)).collate(json.loads(json.dumps(questions[lo:hi])))
    interp = orc.OracleInterpreter(ont, p)
    results = [interp.run(pb, True) for pb in pbs]
    loss = orc.compute_loss(results, [pb._answers for pb in pbs]) / total
    loss.backward()
    return p, loss.detach()


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    ont, questions, counts, feats, bidx, params = _setup()
    lo, hi = shard_range(len(questions), rank, world)
    assert shard_questions(questions, rank, world) == questions[lo:hi]
    p, loss = _grads(ont, params, questions, counts, feats, bidx, lo, hi, len(questions))
    keys = sorted(p)
    bucket = FlatBucket([p[k] for k in keys])
    for k in keys:
        if p[k].grad is not None:
            bucket.grads[id(p[k])].copy_(p[k].grad)
    bucket.all_reduce()
    dist.all_reduce(loss)
    if rank == 0:
        torch.save({'flat_grad': bucket.flat_grad.clone(), 'loss': loss, 'keys': keys}, out)
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_equals_full_batch(tmp_path):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    out = str(tmp_path / 'dp.pt')
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    ont, questions, counts, feats, bidx, params = _setup()
    p, loss = _grads(ont, params, questions, counts, feats, bidx, 0, len(questions), len(questions))
    ref = torch.cat([(p[k].grad if p[k].grad is not None else torch.zeros_like(p[k])).reshape(-1) for k in got['keys']])
    assert torch.allclose(got['flat_grad'], ref, rtol=1e-5, atol=1e-7)
    assert abs(float(got['loss']) - float(loss)) <= 1e-6 * max(1.0, abs(float(loss)))


def test_shard_ranges_cover_everything():
    for total in (1, 7, 256, 4096):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_compiler_relation_slots_are_consistent():
    """Demand-driven relation slots: per image the slot numbering is dense, every relate / choose_rel operand and its
    gradient slice point at the slot of the relation the dense compilation names, and W rows are concept rows."""
    import numpy as np
    from dfol_vqa_b200 import synth
    from dfol_vqa_b200.capi import K
    from dfol_vqa_b200.compiler import ProgramCompiler
    from dfol_vqa_b200.ontology import synthetic_ontology
    from dfol_vqa_b200.programs import ProgramCollater
    ont = synthetic_ontology(160, 24, 5, 4, seed=1, embedding_dim=16)
    for terminal in ('verify_rel', 'choose_rel', 'exist'):
        questions = synth.make_questions(ont, 12, terminal, 1, 3, seed=7, relate_prob=0.7)
        counts = synth.object_counts(12, 9, True, seed=8)
        pb = ProgramCollater(1, lambda qs: (None, None)).collate(questions)[0]
        dense = ProgramCompiler(ont, relation_slots=False).compile(pb, counts)
        slot = ProgramCompiler(ont, relation_slots=True).compile(pb, counts)
        assert dense.img_slot is None and slot.img_slot is not None
        assert dense.instr.shape == slot.instr.shape and dense.lp_num == slot.lp_num
        n_slots = np.diff(slot.img_slot)
        rel_index = list(ont._relation_index)
        for q in range(12):
            seen = {}
            for ip in range(dense.q_instr[q], dense.q_instr[q + 1]):
                d, s = dense.instr[ip], slot.instr[ip]
                assert d[0] == s[0]
                if d[0] == K.OP_RELATE:
                    seen.setdefault(int(d[2]), int(s[2]))
                    assert seen[int(d[2])] == int(s[2])
                elif d[0] == K.OP_CHOOSE_REL:
                    for k in range(int(d[3])):
                        dw, sw = int(dense.opts[d[2] + k]), int(slot.opts[s[2] + k])
                        assert (dw & K.OPT_NEG) == (sw & K.OPT_NEG)
                        seen.setdefault(dw & ~K.OPT_NEG, sw & ~K.OPT_NEG)
                        assert seen[dw & ~K.OPT_NEG] == sw & ~K.OPT_NEG
            assert sorted(seen.values()) == list(range(int(n_slots[q])))
            for col, sl in seen.items():
                assert int(slot.slot_wrow[slot.img_slot[q] + sl]) == rel_index[col]
        assert len(dense.rel_slices) == len(slot.rel_slices)
        for a, b in zip(dense.rel_slices, slot.rel_slices):
            assert a[0] == b[0] and a[2] == b[2] and b[3] == rel_index[a[1]]
        stride = [(c * c + 3) // 4 * 4 for c in counts]
        assert slot.rel_slot_size == max(1, int(sum(k * s for k, s in zip(n_slots, stride))))


def test_collater_attaches_compiled_programs():
    """ProgramCollater(compiler=...) lowers the programs at collate time (DataLoader worker) and the result pickles."""
    import pickle
    from dfol_vqa_b200 import synth
    from dfol_vqa_b200.compiler import ProgramCompiler
    from dfol_vqa_b200.ontology import synthetic_ontology
    from dfol_vqa_b200.programs import ProgramCollater
    ont = synthetic_ontology(160, 24, 5, 4, seed=1, embedding_dim=16)
    questions = synth.make_questions(ont, 10, 'verify_rel', 1, 3, seed=7, relate_prob=0.7)
    counts = synth.object_counts(10, 9, True, seed=8)
    feats, bidx = synth.make_object_features(counts, 32, seed=9)
    comp = ProgramCompiler(ont, relation_slots=True)
    pb = ProgramCollater(1, lambda qs: (feats, bidx), compiler=comp).collate(questions)[0]
    assert pb._dfol_counts == counts
    cp = next(iter(pb._dfol_compiled.values()))
    ref = comp.compile(pb, counts)
    assert (cp.instr == ref.instr).all() and (cp.q_instr == ref.q_instr).all() and cp.rel_slices == ref.rel_slices
    clone = pickle.loads(pickle.dumps(pb))
    assert (next(iter(clone._dfol_compiled.values())).instr == cp.instr).all()
