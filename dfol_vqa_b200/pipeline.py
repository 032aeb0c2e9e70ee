"""Host-side staging pipeline of the step: pinned host batches -> device, double buffered.

The reference trainer moves every program batch to the GPU inside the step (VQATrainer._run_model ->
ProgramBatch.to_cuda, reference src/nsvqa/train/trainer.py:90-97, data_pipeline.py:116-140) and reads the loss back
with .item() right after (trainer.py:440-447), so copy, compute and read-back serialise.  Here the H2D copy of batch
i+1 runs on a copy stream while batch i computes, and the result of step i is read back (pinned, asynchronous)
after step i+1 has been enqueued; every step still pays its own H2D copy and its own D2H read.
"""

import torch


class HostStepPipeline(object):

    def __init__(self, step_fn, device, cold=False):
        """``step_fn(device_program_batch) -> device tensor`` (loss scalar or log-probabilities).  ``cold``: forget the
        device copies of a batch's program tables / targets before staging it, as if every batch were new (a benchmark
        that cycles through a small pool of batches would otherwise find them cached on the device)."""
        self.step_fn = step_fn
        self.cold = cold
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self._host = [None, None]

    def _stage(self, hb):
        if self.cold:
            for cp in getattr(hb, '_dfol_compiled', {}).values():
                cp.device_cache = None
                cp.mod_cache.clear()
            if hasattr(hb, '_dfol_targets'):
                del hb._dfol_targets
        with torch.cuda.stream(self.copy_stream):
            db = hb.to_cuda(self.device.index, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return db, ev

    def run(self, host_batches):
        """Runs the step over ``host_batches`` (any iterable, e.g. a DataLoader that lowers the programs in worker
        processes: it is pulled one batch ahead of the compute); returns the list of results as CPU tensors."""
        compute = torch.cuda.current_stream(self.device)
        it = iter(host_batches)
        results, pending = [], None

        def stage_next():
            hb = next(it, None)
            return None if hb is None else self._stage(hb)

        nxt = stage_next()
        i = 0
        while nxt is not None:
            db, ev = nxt
            nxt = stage_next()
            compute.wait_event(ev)
            # everything to_cuda allocated on the copy stream (features, batch index, staged tensors, packed program
            # tables, loss targets) is consumed on the compute stream: without record_stream the caching allocator
            # could hand a block back to the copy stream -- and the staging of batch i+2 overwrite it -- while step i
            # has not run yet
            for t in db._dfol_device_tensors:
                t.record_stream(compute)
            out = self.step_fn(db).detach()
            slot = i & 1
            i += 1
            if self._host[slot] is None or self._host[slot].shape != out.shape or self._host[slot].dtype != out.dtype:
                self._host[slot] = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
            self._host[slot].copy_(out, non_blocking=True)
            rev = torch.cuda.Event()
            rev.record(compute)
            if pending is not None:
                pending[1].synchronize()
                results.append(pending[0].clone())
            pending = (self._host[slot], rev)
        if pending is not None:
            pending[1].synchronize()
            results.append(pending[0].clone())
        return results
