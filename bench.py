#!/usr/bin/env python
"""Benchmark of the reasoning hot path (BASELINE.json metric: questions/sec on synthetic GQA-shaped data; logic-op %
of the HBM roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3|c4|c0]
                  [--mode train|infer] [--gemm fp32|bf16] [--only]

A *step* is one pass of the hot path over one batch of B questions per GPU:
  train : scene build (featurizer + attribute/relation tables) + program forward + loss + backward + gradient
          all-reduce (N > 1) + clip + Adam      == VQATrainer._train_batch (reference trainer.py:429-442)
  infer : scene build + program forward (+ answers read back in the e2e leg)
N > 1 is launched by torchrun (one rank per GPU, NCCL); questions are sharded by rank (independent scene graphs, no
data-path collective), weak scaling: every rank runs the same per-GPU workload.

Default workload (BASELINE.json configs): c4 at every N -- the data-parallel full-curriculum mixture the metric is quoted
on "at 1/2/4/8 B200" (512 questions per GPU: global batch 4096 at 8 GPUs, all 13 terminal types, one type per batch as
the reference sampler draws them, data_pipeline.py:808-820), so that the per-N lines are the same workload.  At N = 1
the line also carries the complete results (value, e2e, roofline, logic_roofline, kernel table) of the single-GPU
configurations c1, c2 and c3 (the largest: N = 100, nine relate hops) under ``configs``.  ``--workload X --only`` times
one workload alone.

Prints ONE JSON line (rank 0):
  value        questions/s with the inputs resident in HBM (CUDA events around exactly K steps, max over ranks)
  e2e          the same step through the public API from pinned HOST batches: H2D copy of the box features + packed
               program tables and D2H read of the result inside the timed region, every step; reported next to the
               measured pinned-host -> device copy ceiling of the slowest rank (all ranks copying at once)
  e2e_with_lowering  the same with the programs of every step ALSO lowered to bytecode inside the timed region, by
               DataLoader worker processes (ProgramCollater(compiler=...)): bound by host cores per rank
  roofline     dominant kernel of the step, per-launch CUDA-event timings recorded live; tensor AND hbm fractions,
               ``bound`` as SURVEY.md 8(d) classifies the kernel (oracle GEMMs: tensor; logic ops: hbm)
  logic_roofline  the interpreter kernels (program_fwd / program_bwd): algorithmic bytes / time / measured HBM peak
  sustained    the device-resident leg repeated back to back for >= 2 s (clocks under sustained load)
  cpu_baseline the CPU oracle port timed on this box's host cores (bounded sample)
``--impl reference`` times that CPU port alone (the Python reference cannot travel to the GPU box).
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

ALL_TERMINALS = ('exist', 'and', 'or', 'verify_attrs', 'verify_rel', 'all_same', 'all_different', 'two_same',
                 'two_different', 'choose_attr', 'choose_rel', 'query_attr', 'compare')
WORKLOADS = {
    # name: batch per GPU, objects, terminals (one type per batch, as the reference sampler does), hops, relate prob
    'c1': dict(batch=256, n=48, terminals=('exist', 'verify_attrs', 'verify_rel', 'and', 'or'), hops=(1, 3),
               relate_prob=0.35, desc='curriculum stage-1 shape: binary questions, B=256/GPU, N=48, <=3 hops'),
    # (neg_prob: at the trained-like operating point p(attribute) ~ 0.02, so six positive filters in a row drive every
    # attention to the 1e-20 clamp and the whole batch degenerates to constants with zero gradient; mostly-negated
    # filters keep the long chains in the non-degenerate regime -- same instructions, same bytes, meaningful numbers)
    'c2': dict(batch=512, n=48, terminals=('query_attr', 'choose_attr'), hops=(6, 8), relate_prob=0.3, neg_prob=0.85,
               desc='open questions (query/choose over attribute categories), B=512/GPU, N=48, >=6 hops'),
    'c3': dict(batch=256, n=100, terminals=('chain9',), hops=(9, 9), relate_prob=1.0,
               desc='relation-heavy long programs: N=100, 9 relate hops, B=256/GPU'),
    'c4': dict(batch=512, n=48, terminals=ALL_TERMINALS, hops=(1, 9), relate_prob=0.35,
               desc='full-curriculum mixture (all 13 terminal types, one per batch), 1-9 hops, N=48, B=512/GPU '
                    '(4096 global at 8 GPUs)'),
    'c0': dict(batch=32, n=48, terminals=('exist', 'verify_rel', 'and'), hops=(1, 3), relate_prob=0.35,
               desc='sample_config CPU case: B=32, N=48, binary 1-3-hop programs'),
}
DIMS = dict(box=2048, feat=512, hidden=256, emb=300)
VOCAB = dict(concept_num=2335, relation_num=333, category_num=31, class_num=53)
# trained-like operating point (most concept probabilities near 0): keeps the quantifiers out of saturation so that
# losses and gradients of the synthetic workload are finite and meaningful
EMB_BIAS = -4.0


def make_workload_questions(ont, wl, batch, seed, index=0):
    """Questions of batch number ``index`` of a workload (one terminal type per batch)."""
    from dfol_vqa_b200 import synth
    term = wl['terminals'][index % len(wl['terminals'])]
    if term == 'chain9':
        return synth.make_relation_chain_questions(ont, batch, 9, seed=seed)
    return synth.make_questions(ont, batch, term, wl['hops'][0], wl['hops'][1], seed=seed,
                                relate_prob=wl['relate_prob'], neg_prob=wl.get('neg_prob', 0.2))


def build_model(args, device):
    import torch
    from dfol_vqa_b200.factory import build_interpreter
    from dfol_vqa_b200.ontology import synthetic_ontology
    ont = synthetic_ontology(seed=1, embedding_dim=DIMS['emb'], **VOCAB)
    interp = None
    if device is not None:
        extra = {}
        if args.calibrate:
            # sample_config.yaml's arrangement: the four oracle networks frozen, dropout 0.1, the attention-transfer
            # calibrator (two LSTMCells + output layer, state 50) training
            from dfol_vqa_b200.networks import build_attention_networks
            nets = build_attention_networks(DIMS['emb'], 50)
            torch.manual_seed(3)
            with torch.no_grad():  # away from the identity initialisation, as after some training
                nets['attention_output_network'][0].weight.normal_(0.0, 0.05)
            extra = dict(attention_nets=[nets[k] for k in ('forward_attention_network', 'backward_attention_network',
                                                           'attention_output_network')],
                         freeze_oracle=True, dropout=args.dropout)
        if args.train_dropout > 0 and not args.calibrate:
            extra = dict(dropout=args.train_dropout)   # trainable oracle networks under dropout (both passes masked)
        interp = build_interpreter(ont, DIMS, seed=0, device=device, gemm_mode=args.gemm, emb_bias=EMB_BIAS, **extra)
    return ont, interp


def build_batches(ont, wl, B, rank, pool):
    """``pool`` host program batches of the workload (+ the question lists they were collated from)."""
    from dfol_vqa_b200 import synth
    from dfol_vqa_b200.programs import ProgramCollater
    batches, questions = [], []
    for i in range(pool):
        seed = 1000 * rank + i
        qs = make_workload_questions(ont, wl, B, seed, index=i)
        feats, bidx = synth.make_object_features([wl['n']] * B, DIMS['box'], seed=seed + 7)
        pb = ProgramCollater(1, lambda q, f=feats, b=bidx: (f, b)).collate(json.loads(json.dumps(qs)))[0]
        batches.append(pb)
        questions.append(qs)
    return batches, questions


def algorithmic_flops(B, n, C, nR):
    """Forward FLOPs of the scene build per batch (with the pair first layer evaluated through the
    U/V decomposition: 2*T*(F+4)*2H + 8 FLOP per pair and hidden unit instead of 2*P*(2F+12)*H)."""
    T, P = B * n, B * n * n
    F, H, E, D = DIMS['feat'], DIMS['hidden'], DIMS['emb'], DIMS['box']
    feat = 2.0 * T * D * F
    attr = 2.0 * T * ((F + 4) * H + H * E + E * C)
    rel = 2.0 * T * (F + 4) * 2 * H + 2.0 * P * (H * E + E * nR)
    return feat + attr + rel


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.proc, self.lines, self.index = None, [], index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm, mx, power, reasons = [], None, [], set()
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'),
                                 parts[4:8]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        # median over the samples taken under load (upper half: the idle tail is excluded)
        loaded = sm[len(sm) // 2:] if sm else []
        med = loaded[len(loaded) // 2] if loaded else None
        return {'sm_mhz': med, 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm),
                'power_w_max': max(power) if power else None}


def cpu_baseline(args, workload, seconds=12.0):
    """The CPU oracle port on the host cores (bounded sample of the workload).  The fp32 CPU path can produce a
    probability one ulp above 1 on an unlucky sample (binary_cross_entropy then refuses it, exactly as the reference's
    own loss would, trainer.py:196): such a sample is re-drawn, up to three times."""
    last = None
    for attempt in range(3):
        try:
            return _cpu_baseline_once(args, workload, seconds, question_seed=5 + 101 * attempt)
        except RuntimeError as exc:
            last = exc
    raise last


def _cpu_baseline_once(args, workload, seconds, question_seed):
    """The CPU oracle port (oracle/dfol_oracle.py) on the host cores: full train (or infer) step -- collation included,
    as in the reference's step -- on a bounded sample of the same workload."""
    import torch
    sys.path.insert(0, os.path.join(REPO, 'oracle'))
    import dfol_oracle as orc
    from dfol_vqa_b200 import synth
    from dfol_vqa_b200.factory import model_config
    from dfol_vqa_b200.networks import build_networks
    from dfol_vqa_b200.ontology import synthetic_ontology
    from dfol_vqa_b200.programs import ProgramCollater
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = WORKLOADS[workload]
    ont = synthetic_ontology(seed=1, embedding_dim=DIMS['emb'], **VOCAB)
    torch.manual_seed(0)
    nets = build_networks(model_config(DIMS), ont)
    nets['embedding_network']._network[1].bias.data.fill_(EMB_BIAS)
    names = {'featurizer_network': '_featurizer._featurizer_network', 'attribute_network': '_oracle._attribute_network',
             'relation_network': '_oracle._relation_network', 'embedding_network': '_oracle._embedding_network'}
    params = {}
    for key, prefix in names.items():
        for k, v in nets[key].state_dict().items():
            params[prefix + '.' + k] = v.clone().requires_grad_(args.mode == 'train')
    sample_b = args.cpu_sample
    qs = make_workload_questions(ont, wl, sample_b, question_seed, index=0)
    feats, bidx = synth.make_object_features([wl['n']] * sample_b, DIMS['box'], seed=6)

    def one_step():
        pbs = ProgramCollater(1, lambda q: (feats, bidx)).collate(json.loads(json.dumps(qs)))
        if args.mode == 'train':
            for p in params.values():
                p.grad = None
            _, loss = orc.run_step(ont, params, pbs, is_training=True)
            loss.backward()
        else:
            with torch.no_grad():
                orc.run_step(ont, params, pbs, is_training=False)

    one_step()  # warm-up
    times = []
    t_end = time.time() + seconds
    while len(times) < 2 or (time.time() < t_end and len(times) < 50):
        t0 = time.time()
        one_step()
        times.append(time.time() - t0)
    best, med = min(times), sorted(times)[len(times) // 2]
    return {'value': sample_b / med, 'unit': 'questions/s', 'cores': cores, 'kind': 'port',
            'sample': '%d questions of workload %s per step (%s step incl. collation, oracle/dfol_oracle.py = CPU port '
                      'of the reference path, torch CPU fp32 on %d threads, %d timed steps; value = median, best = %.1f '
                      'questions/s)' % (sample_b, workload, args.mode, cores, len(times), sample_b / best),
            'value_best': sample_b / best, 'ms_per_step': med * 1e3}


def default_workload(args):
    """c4 at every N: BASELINE.json quotes its metric 'at 1/2/4/8 B200' on configs[4] (the data-parallel full-curriculum
    mixture, 512 questions per GPU = 4096 global at 8), and the driver computes the scaling efficiency from the per-N
    values of THIS line -- they have to be the same workload.  The single-GPU configurations c1, c2 and c3 (the largest)
    ride along in full under ``configs`` at N = 1."""
    return args.workload or 'c4'


def run_reference(args):
    """--impl reference: the CPU port of the reference path, timed alone (rank 0 only) on this arm's headline config."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    workload = default_workload(args)
    wl = WORKLOADS[workload]
    base = cpu_baseline(args, workload, seconds=max(5.0, 4.0 * args.steps))
    line = {'impl': 'reference', 'metric': 'questions/sec', 'value': base['value'], 'unit': 'questions/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': base['ms_per_step'],
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': '%s: %s' % (workload, wl['desc']), 'step': args.mode,
                       'note': 'bounded CPU sample of the same workload (the CPU path needs O(A T^2) memory per '
                               'program batch: SURVEY.md 8d)'},
            'cpu_baseline': {k: base[k] for k in ('value', 'unit', 'cores', 'kind', 'sample', 'value_best')},
            'e2e': {'value': base['value'], 'unit': 'questions/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


class LoweringDataset(object):
    """Map-style dataset of the e2e leg: item i = the question list of step i; the DataLoader worker that fetches it
    aligns the programs into op slots and lowers them to bytecode + packed tables (ProgramCollater(compiler=...)):
    the collate-time work of the reference's DataLoader workers (data_pipeline.py:893-898).  The box features are NOT
    moved through the workers (reading them from HDF5 is the reference's IO, out of scope): the worker only needs the
    object count of every image; the main process attaches the step's pinned feature tensor."""

    def __init__(self, question_lists, counts, compiler, give_answer, steps):
        import numpy as np
        import torch
        self.q, self.compiler, self.give_answer, self.steps = question_lists, compiler, give_answer, steps
        self.bidx = [torch.from_numpy(np.repeat(np.arange(len(c), dtype=np.int64), c)) for c in counts]

    def __len__(self):
        return self.steps

    def __getitem__(self, i):
        import torch
        from dfol_vqa_b200.programs import ProgramCollater, attach_compiled
        k = i % len(self.q)
        pb = ProgramCollater(1, lambda q: (None, self.bidx[k])).collate(json.loads(self.q[k]))[0]
        attach_compiled(pb, self.compiler, give_answer=self.give_answer)
        if not self.give_answer and pb._answers is not None:
            from dfol_vqa_b200.interpreter import targets_of
            cp = next(iter(pb._dfol_compiled.values()))
            pb._dfol_targets_host = torch.from_numpy(targets_of(cp, pb._answers))
            # the training step reads the packed tables and the targets only: the op-slot objects (argument strings)
            # and the option lists stay in the worker instead of being pickled to the trainer every step
            pb.strip_for_training()
        return k, pb


def h2d_ceiling(device, nbytes=256 << 20, bufs=4, reps=8):
    """Pinned-host -> device copy bandwidth of THIS rank while all ranks copy at the same time (GB/s).  The source
    rotates over ``bufs`` distinct pinned buffers (1 GiB in all, far beyond the host's last-level cache): a single
    128 MiB buffer copied repeatedly is partly served from the CPU cache and overstates what a stream of fresh batches
    can reach (measured at 8 ranks: 37 GB/s per rank against 23 GB/s of the e2e leg's own copies)."""
    import torch
    srcs = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(bufs)]
    for s_ in srcs:
        s_.fill_(1)          # touch every page (first-touch NUMA placement, no lazily mapped zero pages)
    dst = torch.empty(nbytes, dtype=torch.uint8, device=device)
    dst.copy_(srcs[0], non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        dst.copy_(srcs[i % bufs], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


def _gpu_numa_node(index):
    import torch
    pr = torch.cuda.get_device_properties(index)
    path = '/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
    return int(open(path).read())


def _node_cores(node):
    out = []
    for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
        a, _, b = part.partition('-')
        out += list(range(int(a), int(b or a) + 1))
    return out


def pin_rank_to_local_cores(local_rank, local_world):
    """Give every rank its own slice of the host cores (its DataLoader workers inherit it and its pinned staging
    buffers are first-touched from it), on the NUMA node of its GPU when the topology is readable.  Without this the N
    ranks of a box schedule on the same cores and their pinned buffers land wherever the first thread ran."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        return os.cpu_count() or 1, 'affinity unavailable'
    share, how = None, 'even split of the allowed cores'
    try:
        nodes = [_gpu_numa_node(i) for i in range(local_world)]
        mine = nodes[local_rank]
        if mine >= 0:
            cores = [c for c in _node_cores(mine) if c in allowed]
            peers = [r for r in range(local_world) if nodes[r] == mine]
            k = peers.index(local_rank)
            cand = cores[k * len(cores) // len(peers):(k + 1) * len(cores) // len(peers)]
            if cand:
                share, how = cand, 'NUMA node %d of the GPU, shared by %d ranks' % (mine, len(peers))
    except Exception:
        share = None
    if share is None:
        share = allowed[local_rank * len(allowed) // local_world:(local_rank + 1) * len(allowed) // local_world] or allowed
    try:
        os.sched_setaffinity(0, share)
    except OSError:
        return len(allowed), 'sched_setaffinity refused'
    return len(share), how


class Bench(object):

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.args = args
        self.rank = int(os.environ.get('RANK', '0'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        if not torch.cuda.is_available():
            raise SystemExit('bench.py needs a CUDA device: dfol_vqa_b200 has no CPU fallback')
        self.my_cores, self.affinity = pin_rank_to_local_cores(self.local_rank,
                                                                int(os.environ.get('LOCAL_WORLD_SIZE', self.world)))
        torch.set_num_threads(max(1, min(4, self.my_cores)))
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device('cuda', self.local_rank)
        self.group = None
        if self.world > 1:
            # NCCL prints its version banner on fd 1 when the communicator is created: keep stdout to the ONE JSON line
            sys.stdout.flush()
            saved_fd = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group('nccl', device_id=self.device)
                self.group = dist.group.WORLD
                warm = torch.zeros(1, device=self.device)
                dist.all_reduce(warm)
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved_fd, 1)
                os.close(saved_fd)
        self.ont, self.interp = build_model(args, self.device)
        from dfol_vqa_b200.interpreter import FusedTrainStep
        self.trainer = FusedTrainStep(self.interp, process_group=self.group) if args.mode == 'train' else None
        self.interp.train(args.mode == 'train')
        try:
            self.peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
        except (OSError, ValueError):
            self.peaks = {}
        self.traffic = {}
        for name in ('r2_traffic.json', 'r1_traffic.json'):
            try:
                for k, v in json.load(open(os.path.join(REPO, 'profiles', name))).items():
                    self.traffic.setdefault(k, v)
            except (OSError, ValueError):
                pass

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            t = torch.tensor([x], device=self.device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t)
        return x

    def min_over_ranks(self, x):
        return -self.max_over_ranks(-x)

    # ------------------------------------------------------------------------------------------------------

    def run_workload(self, workload, steps, warmup, headline):
        import torch
        from dfol_vqa_b200 import capi
        from dfol_vqa_b200.pipeline import HostStepPipeline
        from dfol_vqa_b200.programs import attach_compiled
        args, interp, trainer, device, world = self.args, self.interp, self.trainer, self.device, self.world
        wl = WORKLOADS[workload]
        B = args.local_batch or wl['batch']
        pool = args.pool or max(5, min(len(wl['terminals']), 13))
        warmup = max(warmup, 3, pool)  # every distinct batch is stepped once before timing (allocator, device tables)
        give_answer = args.mode != 'train'
        host_batches, question_lists = build_batches(self.ont, wl, B, self.rank, pool)
        for pb in host_batches:  # collate-time work: lower the programs to bytecode, pack the tables, pin everything
            attach_compiled(pb, interp._compiler, give_answer=give_answer)
            pb.pin_memory()
        dev_batches = [pb.to_cuda(self.local_rank) for pb in host_batches]
        for db in dev_batches:
            interp.compiled(db, give_answer)
        global_q = B * world

        def step_device(pb):
            if trainer is not None:
                return trainer.step([pb], global_question_num=global_q)
            with torch.no_grad():
                return interp([pb], True)['log_probability']

        pipeline = HostStepPipeline(step_device, device, cold=True)

        def timed(batches, n_steps, n_warm, trace=False, host=False):
            order = [batches[(n_warm + i) % len(batches)] for i in range(n_steps)]
            if host:
                pipeline.run([batches[i % len(batches)] for i in range(n_warm)])
            else:
                for i in range(n_warm):
                    step_device(batches[i % len(batches)])
            self.barrier()
            capi.trace = [] if trace else None
            l0 = capi.launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if host:
                pipeline.run(order)
            else:
                for pb in order:
                    step_device(pb)
            e1.record()
            self.barrier()
            ms = e0.elapsed_time(e1)
            tr, capi.trace = capi.trace, None
            return self.max_over_ranks(ms), capi.launches - l0, tr

        # ---- device-resident leg (value) with the clocks sampled during it
        sampler = ClockSampler(self.local_rank).start() if self.rank == 0 else None
        ms, launches, _ = timed(dev_batches, steps, warmup)
        clocks = sampler.stop() if sampler is not None else None

        # ---- sustained leg: the same step back to back for >= args.sustain seconds
        sustained = None
        if headline and args.sustain > 0:
            n_sus = max(steps, int(args.sustain * 1e3 / max(ms / steps, 1e-3)) + 1)
            sampler = ClockSampler(self.local_rank).start() if self.rank == 0 else None
            ms_sus, _, _ = timed(dev_batches, n_sus, 1)
            cl = sampler.stop() if sampler is not None else None
            sustained = {'seconds': ms_sus * 1e-3, 'steps': n_sus, 'ms_per_step': ms_sus / n_sus,
                         'value': global_q * n_sus / (ms_sus * 1e-3), 'unit': 'questions/s', 'clocks': cl}

        if args.no_e2e:
            ms_pre = ms
            e2e = {'value': None, 'ms_per_step': None, 'h2d_ceiling_gbs': 1.0, 'h2d_ceiling_note': 'skipped (--no-e2e)'}
        else:
            # ---- e2e: pinned host batches, H2D of features + tables and D2H of the result per step
            ms_pre, _, _ = timed(host_batches, steps, warmup, host=True)
            # ---- e2e with the lowering inside: DataLoader workers collate + compile the programs of every step
            e2e = self.e2e_with_lowering(host_batches, question_lists, pipeline, wl, B, steps, warmup, give_answer)

        ms_staged = staged_bytes = None
        if args.gemm == 'bf16' and headline and not args.no_e2e:
            # same pre-lowered leg with the collate-time bf16 staging of the box features (ProgramBatch.stage_bf16):
            # the device casts them to bf16 as its first step anyway, bit-identical results, half the H2D bytes
            import copy
            staged = [copy.copy(pb).stage_bf16(drop_fp32=True).pin_memory() for pb in host_batches]
            ms_staged, _, _ = timed(staged, steps, warmup, host=True)
            staged_bytes = int(sum(t.numel() * t.element_size() for t in staged[0]._staged) +
                               staged[0]._object_batch_index.numel() * 8 +
                               sum(cp.blob.numel() for cp in staged[0]._dfol_compiled.values()))
            del staged

        # ---- per-kernel pass with CUDA events around every launch (same steps, same stream)
        ms_tr, _, tr = timed(dev_batches, steps, 1, trace=True)

        table_bytes = int(sum(cp.blob.numel() for cp in host_batches[0]._dfol_compiled.values()) +
                          getattr(host_batches[0], '_dfol_targets_host', torch.zeros(0)).numel() * 4)
        feat_bytes = int(host_batches[0]._object_features.numel() * 4 +
                         host_batches[0]._object_batch_index.numel() * 8) + table_bytes
        pair_rows = [next(iter(pb._dfol_compiled.values())).layout_meta for pb in host_batches]
        del dev_batches, host_batches
        torch.cuda.empty_cache()
        if self.rank != 0:
            return None

        per = {}
        for name, meta, a, b in tr:
            key = meta.get('tag') or name
            d = per.setdefault(key, {'ms': 0.0, 'n': 0, 'flops': 0.0, 'bytes': 0.0, 'entry': name})
            d['ms'] += a.elapsed_time(b)
            d['n'] += 1
            d['flops'] += meta.get('flops', 0.0)
            d['bytes'] += meta.get('bytes', 0.0)
        total_kernel_ms = sum(d['ms'] for d in per.values()) or 1.0
        t_peak = self.peaks.get('bf16_tflops_sustained', 1400.0)
        h_peak = self.peaks.get('hbm_gbs', 6650.0)

        def roof_of(key):
            d = per[key]
            t_ach = d['flops'] / (d['ms'] * 1e-3) / 1e12
            h_ach = d['bytes'] / (d['ms'] * 1e-3) / 1e9
            # the oracle's GEMMs (featurizer, attribute chain, pair-level layers and their dgrad / wgrad)
            # are reported against the tensor roofline; everything else on the path -- the logic ops, the pair hidden
            # layer, the table-layer backward (a K = 16..32 contraction: 24-40 FLOP/B, far below the ~210 FLOP/B ridge,
            # whichever unit evaluates it) -- against the HBM roofline.  Both fractions are always printed.
            tensor = d['flops'] > 0 and key.startswith(('gemm_', 'pair_layer_', 'pair_chain'))
            roof = {'kernel': key, 'entry_point': d['entry'], 'bound': 'tensor' if tensor else 'hbm',
                    'achieved': t_ach if tensor else h_ach, 'peak': t_peak if tensor else h_peak,
                    'unit': 'TFLOP/s' if tensor else 'GB/s', 'frac': (t_ach / t_peak) if tensor else (h_ach / h_peak),
                    'tensor_frac': t_ach / t_peak if d['flops'] else None,
                    'hbm_frac': h_ach / h_peak if d['bytes'] else None,
                    'algorithmic_flops_per_launch': d['flops'] / max(d['n'], 1),
                    'algorithmic_bytes_per_launch': d['bytes'] / max(d['n'], 1),
                    'traffic': self.traffic.get('%s@%s' % (key, workload)),
                    'launches_per_step': d['n'] / steps, 'avg_launch_ms': d['ms'] / max(d['n'], 1),
                    'share_of_kernel_time': d['ms'] / total_kernel_ms,
                    'peak_source': 'MEASURED_PEAKS.json' if self.peaks else 'fallback (B200_PROFILING.md)'}
            return roof

        top_key = max(per, key=lambda k: per[k]['ms'])
        logic = {}
        for key in ('program_fwd', 'program_bwd'):
            if key in per and per[key]['ms'] > 0:
                d = per[key]
                gbs = d['bytes'] / (d['ms'] * 1e-3) / 1e9
                logic[key] = {'achieved': gbs, 'peak': h_peak, 'unit': 'GB/s', 'frac': gbs / h_peak,
                              'avg_launch_ms': d['ms'] / d['n'],
                              'algorithmic_bytes_per_launch': d['bytes'] / d['n']}
        kernels = {k: {'ms_per_step': v['ms'] / steps, 'launches_per_step': v['n'] / steps,
                       'tflops': (v['flops'] / (v['ms'] * 1e-3) / 1e12) if v['flops'] and v['ms'] else None,
                       'gbs': (v['bytes'] / (v['ms'] * 1e-3) / 1e9) if v['bytes'] and v['ms'] else None}
                   for k, v in sorted(per.items(), key=lambda kv: -kv[1]['ms'])[:40]}
        C, nR = VOCAB['concept_num'], VOCAB['relation_num']
        res = {
            'value': global_q * steps / (ms * 1e-3), 'unit': 'questions/s', 'ms_per_step': ms / steps,
            'steps': steps, 'warmup': warmup,
            'config': {'workload': '%s: %s' % (workload, wl['desc']),
                       'step': args.mode + (' (calibrator only: frozen oracle, dropout %.2f)' % args.dropout
                                            if args.calibrate else '') +
                               (' (oracle networks trained under dropout %.2f)' % args.train_dropout
                                if args.train_dropout > 0 and not args.calibrate else ''), 'gemm_mode': args.gemm,
                       'global_batch': global_q, 'objects_per_image': wl['n'], 'box_feature_dim': DIMS['box'],
                       'concepts': C, 'relations': nR, 'parallelism': 'dp%d (questions sharded by rank)' % world,
                       'pair_rows': 'demand-driven: %d of %d images per batch carry pair rows on average (images whose '
                                    'program reads no relation get none)' % (
                                        sum(m['pair_images'] for m in pair_rows) // len(pair_rows), B),
                       'l2': 'inputs larger than L2: per-step activations + tables %.1f GB >> 126 MB; %d distinct '
                             'batches cycled' % (sum(m['P'] for m in pair_rows) / len(pair_rows) * 1300 / 1e9 +
                                                 feat_bytes / 1e9, pool)},
            'clocks': clocks,
            # e2e (the contract's definition): pinned host batches -> H2D of features + packed tables -> step -> D2H of the
            # result, every step, programs lowered at collate time (outside the timed region)
            'e2e': {'value': global_q * steps / (ms_pre * 1e-3), 'unit': 'questions/s', 'ms_per_step': ms_pre / steps,
                    'h2d_bytes_per_step': feat_bytes, 'd2h_bytes_per_step': 4 if args.mode == 'train' else 4 * B,
                    'h2d_ceiling_gbs': e2e['h2d_ceiling_gbs'], 'h2d_ceiling_note': e2e['h2d_ceiling_note'],
                    'h2d_gbs_achieved': feat_bytes * steps / (ms_pre * 1e-3) / 1e9,
                    'h2d_frac_of_ceiling': feat_bytes * steps / (ms_pre * 1e-3) / 1e9 / e2e['h2d_ceiling_gbs'],
                    'note': 'per-rank copy rate against the measured per-rank ceiling: the e2e step is bound by the host -> '
                            'device copy of the fp32 box features (the reference collator\'s format), not by the GPU'},
            # the same with the program lowering INSIDE the timed region, in DataLoader worker processes
            'e2e_with_lowering': dict({k: v for k, v in e2e.items() if not k.startswith('h2d_ceiling')},
                                      unit='questions/s', h2d_bytes_per_step=feat_bytes),
            'gpu_launches': launches,
            'roofline': roof_of(top_key),
            'logic_roofline': logic,
            'kernels': kernels,
            'kernel_ms_per_step': total_kernel_ms / steps,
            'scene_fwd_gflop_per_step_dense': algorithmic_flops(B, wl['n'], C, nR) / 1e9,
        }
        if sustained is not None:
            res['sustained'] = sustained
        if ms_staged is not None:
            res['e2e_bf16_staging'] = {
                'value': global_q * steps / (ms_staged * 1e-3), 'unit': 'questions/s',
                'h2d_bytes_per_step': staged_bytes, 'ms_per_step': ms_staged / steps,
                'note': 'optional host format (ProgramBatch.stage_bf16: bf16 features + fp32 geometry, cast at collate '
                        'time), programs pre-lowered; bit-identical results in tensor-core mode; NOT the headline e2e'}
        return res

    def e2e_with_lowering(self, host_batches, question_lists, pipeline, wl, B, steps, warmup, give_answer):
        """The e2e leg proper: DataLoader workers lower the programs of every step inside the timed region."""
        import torch
        from torch.utils.data import DataLoader
        args = self.args
        workers = args.workers if args.workers >= 0 else max(1, min(12, self.my_cores - 1))
        counts = [[wl['n']] * B for _ in question_lists]
        n_items = warmup + steps
        ds = LoweringDataset([json.dumps(q) for q in question_lists], counts, self.interp._compiler, give_answer, n_items)
        feats = [(pb._object_features, pb._object_batch_index) for pb in host_batches]

        def attach(item):
            k, pb = item
            pb._object_features, pb._object_batch_index = feats[k]
            pb.pin_memory()   # packed tables + targets (features are pinned already)
            return pb

        # pin_memory: the loader's pin thread (not the training thread) unpickles the workers' results and pins the packed
        # tables / targets through ProgramBatch.pin_memory()
        # The workers are forked: they inherit every live CUDA tensor of this process, and freeing one whose streams were
        # recorded (HostStepPipeline) calls into a CUDA context the child does not have -- a cyclic-GC run in a worker
        # that finds inherited garbage kills it.  Collect the garbage here and freeze what is left out of the
        # children's collector.
        import gc
        gc.collect()
        gc.freeze()
        loader = DataLoader(ds, batch_size=None, shuffle=False, num_workers=workers, pin_memory=True,
                            prefetch_factor=4 if workers > 0 else None, persistent_workers=False)
        it = iter(loader)
        gc.unfreeze()
        pipeline.run([attach(next(it)) for _ in range(warmup)])
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        pipeline.run(attach(item) for item in it)
        e1.record()
        self.barrier()
        wall = time.time() - t0
        ms = self.max_over_ranks(e0.elapsed_time(e1))
        ceiling = self.min_over_ranks(h2d_ceiling(self.device))
        return {'value': B * self.world * steps / (ms * 1e-3), 'ms_per_step': ms / steps,
                'lowering': 'inside the timed region: %d DataLoader worker processes per rank align the programs into op '
                            'slots and lower them to bytecode (%d host cores for this rank: %s)' % (workers, self.my_cores,
                                                                                              self.affinity),
                'h2d_ceiling_gbs': ceiling,
                'h2d_ceiling_note': 'pinned host -> device copy rate of the slowest rank while all %d ranks copy at once '
                                    '(8 copies of 256 MiB from 4 distinct pinned buffers)' % self.world,
                'wall_ms_per_step': wall * 1e3 / steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default=None, choices=sorted(WORKLOADS))
    ap.add_argument('--only', action='store_true', help='time --workload alone (no nested configs)')
    ap.add_argument('--mode', default='train', choices=['train', 'infer'])
    ap.add_argument('--gemm', default=None, choices=['fp32', 'bf16'])
    ap.add_argument('--local-batch', type=int, default=0)
    ap.add_argument('--pool', type=int, default=0, help='distinct pre-collated batches cycled through the steps')
    ap.add_argument('--cpu-sample', type=int, default=8)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--sustain', type=float, default=2.0, help='seconds of the sustained leg (0 = off)')
    ap.add_argument('--no-e2e', action='store_true', help='device-resident legs only (profiling under ncu)')
    ap.add_argument('--workers', type=int, default=-1, help='DataLoader workers of the e2e leg (-1 = by host cores)')
    ap.add_argument('--calibrate', action='store_true',
                    help="sample_config.yaml's training arrangement: frozen oracle networks with dropout, the "
                         'attention-transfer calibrator trains (not the default headline workload)')
    ap.add_argument('--dropout', type=float, default=0.1)
    ap.add_argument('--train-dropout', type=float, default=0.0,
                    help='train the oracle networks with this dropout probability (not the default headline workload)')
    args = ap.parse_args()
    if args.gemm is None:
        args.gemm = 'bf16'  # tensor-core mode (bf16 operands, fp32 accumulation); --gemm fp32 = parity mode
    if args.impl == 'reference':
        return run_reference(args)

    import torch.distributed as dist
    bench = Bench(args)
    headline = default_workload(args)
    nested = [] if (args.only or args.workload or bench.world > 1 or args.calibrate or args.train_dropout > 0
                    or args.gemm != 'bf16') else ['c1', 'c2', 'c3']
    res = bench.run_workload(headline, args.steps, args.warmup, headline=True)
    configs = {}
    for w in nested:
        r = bench.run_workload(w, args.steps, args.warmup, headline=False)
        if r is not None:
            configs[w] = r
    if bench.rank == 0:
        line = {'metric': 'questions/sec', 'n_gpus': bench.world, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32' if args.gemm == 'fp32' else 'bf16', 'data': 'synthetic'}
        line.update(res)
        if configs:
            line['configs'] = configs
        if not args.no_cpu_baseline:
            try:
                line['cpu_baseline'] = {k: v for k, v in cpu_baseline(args, headline).items() if k != 'ms_per_step'}
            except Exception as exc:  # the baseline leg must not lose the GPU numbers
                line['cpu_baseline'] = {'value': None, 'unit': 'questions/s', 'cores': os.cpu_count(), 'kind': 'port',
                                        'sample': 'failed: %r' % (exc,)}
        print(json.dumps(line))
    if bench.world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
