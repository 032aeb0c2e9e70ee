#!/usr/bin/env python
"""Benchmark of the reasoning hot path (BASELINE.json metric: questions/sec on synthetic GQA-shaped data).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3] [--mode train|infer]
                  [--gemm fp32|bf16]

A *step* is one pass of the hot path over one batch of B questions per GPU:
  train : scene build (featurizer + attribute/relation tables) + program forward + loss + backward + gradient
          all-reduce (N > 1) + clip + Adam      == VQATrainer._train_batch (reference trainer.py:429-442)
  infer : scene build + program forward (+ answers read back in the e2e leg)
N > 1 is launched by torchrun (one rank per GPU, NCCL); questions are sharded by rank (independent scene graphs,
no data-path collective), weak scaling: every rank runs the same per-GPU workload.

Prints ONE JSON line (rank 0). ``value`` = questions/s with the inputs resident in HBM; ``e2e`` = the same step
through the public API with HOST (pinned) buffers: H2D copy of the box features inside the timed region and a D2H
read of the loss.  ``roofline`` describes the dominant kernel of the step (per-launch CUDA-event timings recorded
live in the timed region); ``cpu_baseline`` is the CPU oracle port timed on this box's host cores.
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

WORKLOADS = {
    # name: batch per GPU, objects, terminals (one type per batch, as the reference sampler does), hops, relate prob
    'c1': dict(batch=256, n=48, terminals=('exist', 'verify_attrs', 'verify_rel', 'and', 'or'), hops=(1, 3),
               relate_prob=0.35, desc='curriculum stage-1 shape: binary questions, B=256/GPU, N=48, <=3 hops'),
    'c2': dict(batch=512, n=48, terminals=('query_attr', 'choose_attr'), hops=(6, 8), relate_prob=0.3,
               desc='open questions (query/choose over attribute categories), B=512/GPU, N=48, >=6 hops'),
    'c3': dict(batch=256, n=100, terminals=('chain9',), hops=(9, 9), relate_prob=1.0,
               desc='relation-heavy long programs: N=100, 9 relate hops, B=256/GPU'),
    'c0': dict(batch=32, n=48, terminals=('exist', 'verify_rel', 'and'), hops=(1, 3), relate_prob=0.35,
               desc='sample_config CPU case: B=32, N=48, binary 1-3-hop programs'),
}
DIMS = dict(box=2048, feat=512, hidden=256, emb=300)
VOCAB = dict(concept_num=2335, relation_num=333, category_num=31, class_num=53)
# trained-like operating point (most concept probabilities near 0): keeps the quantifiers out of saturation so that
# losses and gradients of the synthetic workload are finite and meaningful
EMB_BIAS = -4.0


def build_world(args, rank, device):
    import torch
    sys.path.insert(0, os.path.join(REPO, 'tests'))
    import helpers
    from dfol_vqa_b200 import synth
    from dfol_vqa_b200.ontology import synthetic_ontology
    from dfol_vqa_b200.programs import ProgramCollater
    wl = WORKLOADS[args.workload]
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
        interp = helpers.build_interpreter(ont, DIMS, seed=0, device=device, gemm_mode=args.gemm, emb_bias=EMB_BIAS,
                                           **extra)
    B = args.local_batch or wl['batch']
    batches = []
    for i in range(args.pool):
        seed = 1000 * rank + i
        term = wl['terminals'][i % len(wl['terminals'])]
        if term == 'chain9':
            qs = synth.make_relation_chain_questions(ont, B, 9, seed=seed)
        else:
            qs = synth.make_questions(ont, B, term, wl['hops'][0], wl['hops'][1], seed=seed,
                                      relate_prob=wl['relate_prob'])
        counts = [wl['n']] * B
        feats, bidx = synth.make_object_features(counts, DIMS['box'], seed=seed + 7)
        pb = ProgramCollater(1, lambda q, f=feats, b=bidx: (f, b)).collate(qs)[0]
        batches.append(pb)
    return ont, interp, batches, B


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

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
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
        return {'sm_mhz': med, 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


def cpu_baseline(args, seconds=12.0):
    """The CPU oracle port (oracle/dfol_oracle.py) on the host cores: full train (or infer) step on a bounded sample
    of the same workload."""
    import torch
    sys.path.insert(0, os.path.join(REPO, 'oracle'))
    sys.path.insert(0, os.path.join(REPO, 'tests'))
    import dfol_oracle as orc
    import helpers
    from dfol_vqa_b200 import synth
    from dfol_vqa_b200.networks import build_networks
    from dfol_vqa_b200.ontology import synthetic_ontology
    from dfol_vqa_b200.programs import ProgramCollater
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = WORKLOADS[args.workload]
    ont = synthetic_ontology(seed=1, embedding_dim=DIMS['emb'], **VOCAB)
    torch.manual_seed(0)
    nets = build_networks(helpers.model_config(DIMS), ont)
    nets['embedding_network']._network[1].bias.data.fill_(EMB_BIAS)
    names = {'featurizer_network': '_featurizer._featurizer_network', 'attribute_network': '_oracle._attribute_network',
             'relation_network': '_oracle._relation_network', 'embedding_network': '_oracle._embedding_network'}
    params = {}
    for key, prefix in names.items():
        for k, v in nets[key].state_dict().items():
            params[prefix + '.' + k] = v.clone().requires_grad_(args.mode == 'train')
    sample_b = args.cpu_sample
    term = wl['terminals'][0]
    if term == 'chain9':
        qs = synth.make_relation_chain_questions(ont, sample_b, 9, seed=5)
    else:
        qs = synth.make_questions(ont, sample_b, term, wl['hops'][0], wl['hops'][1], seed=5,
                                  relate_prob=wl['relate_prob'])
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
    best = min(times)
    return {'value': sample_b / best, 'unit': 'questions/s', 'cores': cores, 'kind': 'port',
            'sample': '%d questions of workload %s per step (%s step, oracle/dfol_oracle.py, torch CPU fp32, '
                      '%d timed steps, best)' % (sample_b, args.workload, args.mode, len(times)),
            'ms_per_step': best * 1e3}


def run_reference(args):
    """--impl reference: the CPU port of the reference path, timed alone (rank 0 only)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    base = cpu_baseline(args, seconds=max(5.0, 4.0 * args.steps))
    line = {'impl': 'reference', 'metric': 'questions/sec', 'value': base['value'], 'unit': 'questions/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': base['ms_per_step'],
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': '%s (%s), %s step, bounded CPU sample' % (args.workload, wl['desc'], args.mode)},
            'cpu_baseline': {k: base[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
            'e2e': {'value': base['value'], 'unit': 'questions/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c1', choices=sorted(WORKLOADS))
    ap.add_argument('--mode', default='train', choices=['train', 'infer'])
    ap.add_argument('--gemm', default=None, choices=['fp32', 'bf16'])
    ap.add_argument('--local-batch', type=int, default=0)
    ap.add_argument('--pool', type=int, default=5, help='distinct pre-collated batches cycled through the steps')
    ap.add_argument('--cpu-sample', type=int, default=8)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--calibrate', action='store_true',
                    help="sample_config.yaml's training arrangement: frozen oracle networks with dropout, the "
                         'attention-transfer calibrator trains (not the default headline workload)')
    ap.add_argument('--dropout', type=float, default=0.1)
    ap.add_argument('--train-dropout', type=float, default=0.0,
                    help='train the oracle networks with this dropout probability (not the default headline workload)')
    args = ap.parse_args()
    if args.gemm is None:
        args.gemm = 'bf16'  # tensor-core mode (bf16 operands, fp32 accumulation); --gemm fp32 = parity mode
    # every distinct batch of the pool is stepped once before timing (allocator growth, per-batch device tables)
    args.warmup = max(args.warmup, 3, args.pool)
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from dfol_vqa_b200 import capi
    from dfol_vqa_b200.interpreter import FusedTrainStep
    from dfol_vqa_b200.pipeline import HostStepPipeline

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: dfol_vqa_b200 has no CPU fallback')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    group = None
    if world > 1:
        # NCCL prints its version banner on fd 1 when the communicator is created: keep stdout to the ONE JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=device)
            group = dist.group.WORLD
            warm = torch.zeros(1, device=device)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    ont, interp, host_batches, B = build_world(args, rank, device)
    wl = WORKLOADS[args.workload]
    C, nR = VOCAB['concept_num'], VOCAB['relation_num']
    from dfol_vqa_b200.programs import attach_compiled
    for pb in host_batches:  # collate-time work: lower the programs to bytecode, pack the tables, pin everything
        attach_compiled(pb, interp._compiler, give_answer=(args.mode != 'train'))
        pb.pin_memory()
    dev_batches = [pb.to_cuda(local_rank) for pb in host_batches]
    for db in dev_batches:
        interp.compiled(db, args.mode != 'train')
    trainer = FusedTrainStep(interp, process_group=group) if args.mode == 'train' else None
    interp.train(args.mode == 'train')
    global_q = B * world

    def step_device(pb):
        if trainer is not None:
            return trainer.step([pb], global_question_num=global_q)
        with torch.no_grad():
            return interp([pb], True)['log_probability']

    # e2e: public API with HOST (pinned) batches; every step copies its features H2D and reads its result back
    # (HostStepPipeline double-buffers the copy of batch i+1 behind the compute of batch i)
    # (cold: the device copies of a batch's program tables and targets are dropped before it is staged, so every step
    # uploads them again with its features -- a training run never sees the same batch twice)
    pipeline = HostStepPipeline(step_device, device, cold=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, batches, steps, warmup, trace=False, host=False):
        order = [batches[(warmup + i) % len(batches)] for i in range(steps)]
        if host:
            pipeline.run([batches[i % len(batches)] for i in range(warmup)])
        else:
            for i in range(warmup):
                fn(batches[i % len(batches)])
        barrier()
        capi.trace = [] if trace else None
        l0 = capi.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if host:
            pipeline.run(order)
        else:
            for pb in order:
                fn(pb)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        tr, capi.trace = capi.trace, None
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, capi.launches - l0, tr

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, launches, _ = timed(step_device, dev_batches, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _, _ = timed(None, host_batches, args.steps, args.warmup, host=True)
    ms_staged = None
    if args.gemm == 'bf16':
        # same e2e leg with the collate-time bf16 staging of the box features (ProgramBatch.stage_bf16): the device
        # casts them to bf16 as its first step anyway, so the results are bit-identical and the H2D copy halves
        import copy
        staged = [copy.copy(pb).stage_bf16(drop_fp32=True).pin_memory() for pb in host_batches]
        ms_staged, _, _ = timed(None, staged, args.steps, args.warmup, host=True)
        staged_bytes = int(sum(t.numel() * t.element_size() for t in staged[0]._staged) +
                           staged[0]._object_batch_index.numel() * 8 +
                           sum(cp.blob.numel() for cp in staged[0]._dfol_compiled.values()))
        del staged
    # per-kernel pass with CUDA events around every launch (same steps, same stream)
    ms_tr, _, tr = timed(step_device, dev_batches, args.steps, 1, trace=True)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    per = {}
    for name, meta, a, b in tr:
        key = meta.get('tag') or name
        d = per.setdefault(key, {'ms': 0.0, 'n': 0, 'flops': 0.0, 'bytes': 0.0, 'entry': name})
        d['ms'] += a.elapsed_time(b)
        d['n'] += 1
        d['flops'] += meta.get('flops', 0.0)
        d['bytes'] += meta.get('bytes', 0.0)
    total_kernel_ms = sum(d['ms'] for d in per.values()) or 1.0
    top_key = max(per, key=lambda k: per[k]['ms'])
    top = per[top_key]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))
    except (OSError, ValueError):
        pass
    # a kernel may carry both algorithmic FLOPs and algorithmic bytes (the pair-level GEMMs stream ~1 GB of bf16
    # activations per launch): it is reported against the roofline it sits closer to
    t_peak, h_peak = peaks.get('bf16_tflops_sustained', 1400.0), peaks.get('hbm_gbs', 6650.0)
    t_ach = top['flops'] / (top['ms'] * 1e-3) / 1e12
    h_ach = top['bytes'] / (top['ms'] * 1e-3) / 1e9
    if t_ach / t_peak >= h_ach / h_peak:
        roof = {'bound': 'tensor', 'achieved': t_ach, 'peak': t_peak, 'unit': 'TFLOP/s', 'frac': t_ach / t_peak}
    else:
        roof = {'bound': 'hbm', 'achieved': h_ach, 'peak': h_peak, 'unit': 'GB/s', 'frac': h_ach / h_peak}
    roof['tensor_frac'], roof['hbm_frac'] = t_ach / t_peak, h_ach / h_peak
    traffic = None
    try:
        table = json.load(open(os.path.join(REPO, 'profiles', 'r1_traffic.json')))
        traffic = table.get(top_key, table.get('%s@%s' % (top_key, args.workload)))
    except (OSError, ValueError):
        pass
    roof['traffic'] = traffic  # ncu dram bytes per launch (profiles/r1_traffic.json), None if not captured
    roof.update({'kernel': top_key, 'entry_point': top['entry'], 'launches_per_step': top['n'] / args.steps,
                 'avg_launch_ms': top['ms'] / max(top['n'], 1), 'share_of_kernel_time': top['ms'] / total_kernel_ms,
                 'peak_source': 'MEASURED_PEAKS.json' if peaks else 'fallback (B200_PROFILING.md)'})
    kernels = {k: {'ms_per_step': v['ms'] / args.steps, 'launches_per_step': v['n'] / args.steps,
                   'tflops': (v['flops'] / (v['ms'] * 1e-3) / 1e12) if v['flops'] and v['ms'] else None,
                   'gbs': (v['bytes'] / (v['ms'] * 1e-3) / 1e9) if v['bytes'] and v['ms'] else None}
               for k, v in sorted(per.items(), key=lambda kv: -kv[1]["ms"])[:40]}

    value = global_q * args.steps / (ms * 1e-3)
    e2e_value = global_q * args.steps / (ms_e2e * 1e-3)
    table_bytes = int(sum(cp.blob.numel() for cp in host_batches[0]._dfol_compiled.values()) +
                      getattr(host_batches[0], '_dfol_targets_host', torch.zeros(0)).numel() * 4)
    feat_bytes = int(host_batches[0]._object_features.numel() * 4 + host_batches[0]._object_batch_index.numel() * 8) + \
        table_bytes
    fwd_flops = algorithmic_flops(B, wl['n'], C, nR)
    line = {
        'metric': 'questions/sec', 'value': value, 'unit': 'questions/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32' if args.gemm == 'fp32' else 'bf16', 'data': 'synthetic',
        'config': {'workload': '%s: %s' % (args.workload, wl['desc']),
                   'step': args.mode + (' (calibrator only: frozen oracle, dropout %.2f)' % args.dropout
                                        if args.calibrate else '') +
                           (' (oracle networks trained under dropout %.2f)' % args.train_dropout
                            if args.train_dropout > 0 and not args.calibrate else ''), 'gemm_mode': args.gemm,
                   'global_batch': global_q, 'objects_per_image': wl['n'], 'box_feature_dim': DIMS['box'],
                   'concepts': C, 'relations': nR, 'parallelism': 'dp%d (questions sharded by rank)' % world,
                   'l2': 'inputs larger than L2: per-step tables + activations %.1f GB >> 126 MB; %d distinct '
                         'batches cycled' % ((B * wl['n'] ** 2 * (nR + 556) * 4) / 1e9, len(dev_batches))},
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': 'questions/s', 'h2d_bytes_per_step': feat_bytes,
                'd2h_bytes_per_step': 4 if args.mode == 'train' else 4 * B, 'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': launches,
        'e2e_bf16_staging': None if ms_staged is None else {
            'value': global_q * args.steps / (ms_staged * 1e-3), 'unit': 'questions/s',
            'h2d_bytes_per_step': staged_bytes, 'ms_per_step': ms_staged / args.steps,
            'note': 'optional host format (ProgramBatch.stage_bf16: bf16 features + fp32 geometry, cast at collate '
                    'time); bit-identical results in tensor-core mode; NOT the headline e2e, which keeps the '
                    "reference collator's fp32 features"},
        'roofline': roof,
        'kernels': kernels,
        'kernel_ms_per_step': total_kernel_ms / args.steps,
        'scene_fwd_gflop_per_step': fwd_flops / 1e9,
    }
    if not args.no_cpu_baseline and world >= 1:
        try:
            line['cpu_baseline'] = {k: v for k, v in cpu_baseline(args).items() if k != 'ms_per_step'}
        except Exception as exc:  # the baseline leg must not lose the GPU numbers
            line['cpu_baseline'] = {'value': None, 'unit': 'questions/s', 'cores': os.cpu_count(), 'kind': 'port',
                                    'sample': 'failed: %r' % (exc,)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
