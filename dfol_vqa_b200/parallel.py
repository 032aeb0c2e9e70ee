"""Data-parallel plumbing of the training step: question sharding and the flat parameter / gradient bucket.

Replaces ProgramDataParallel (reference: src/nsvqa/nn/interpreter/data_parallel.py:54-83: per-step replicate,
scatter of program batches over devices, reduce-add of gradients onto GPU 0) with one process per GPU, persistent
replicas, rank-local question shards and ONE all-reduce of a flat fp32 gradient bucket per step (NCCL on GPUs; the
same code runs over gloo in the CPU tests)."""

import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous [lo, hi) share of ``total`` questions for ``rank`` (sizes differ by at most one)."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_questions(questions, rank, world):
    lo, hi = shard_range(len(questions), rank, world)
    return questions[lo:hi]


class FlatBucket(object):
    """Re-homes a list of parameters into one contiguous fp32 buffer (``flat``) with a matching gradient buffer
    (``flat_grad``); ``grads[id(p)]`` are views of ``flat_grad`` shaped like the parameters."""

    def __init__(self, params):
        self.params = list(params)
        dev = self.params[0].device
        sizes = [p.numel() for p in self.params]
        self.flat = torch.empty(sum(sizes), device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros_like(self.flat)
        self.grads = {}
        off = 0
        for p, n in zip(self.params, sizes):
            self.flat[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + n].view_as(p)
            self.grads[id(p)] = self.flat_grad[off:off + n].view_as(p)
            off += n

    def zero_grad(self):
        self.flat_grad.zero_()

    def all_reduce(self, group=None):
        """Sum the gradient bucket over the ranks (gradients are pre-scaled by 1/global batch by the caller)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=group)
