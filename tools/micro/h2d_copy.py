"""Pinned-host -> device copy ceiling with every rank copying at once (the bound of bench.py's e2e leg).

  python tools/micro/h2d_copy.py                                   # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
      tools/micro/h2d_copy.py                                        # 8 ranks, one per GPU

Prints one JSON line (rank 0): per-rank GB/s (min / mean / max over ranks) for
  rotating : 8 copies of 256 MiB from 4 distinct pinned buffers (what bench.py reports as e2e.h2d_ceiling_gbs),
  single   : 6 copies of ONE 128 MiB buffer (round-1/2 variant: partly served from the CPU's last-level cache, reads high).
Measured on the pool's 8-GPU box: rotating 54.0 GB/s at 1 rank, 25.0 GB/s per rank at 8 ranks; single 55.4 / 37.2.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def single(device, nbytes=128 << 20, reps=6):
    src = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dst = torch.empty(nbytes, dtype=torch.uint8, device=device)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


def main():
    import bench
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=device)
    out = {}
    for name, fn in (('rotating', bench.h2d_ceiling), ('single', single)):
        if world > 1:
            dist.barrier()
        v = torch.tensor([fn(device)], device=device, dtype=torch.float64)
        if world > 1:
            vs = [torch.zeros_like(v) for _ in range(world)]
            dist.all_gather(vs, v)
            v = torch.cat(vs)
        out[name] = {'min': float(v.min()), 'mean': float(v.mean()), 'max': float(v.max())}
    if int(os.environ.get('RANK', '0')) == 0:
        print(json.dumps({'ranks': world, 'unit': 'GB/s per rank', **out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
