"""HBM bandwidth by direction on this GPU: copy (read + write) and write only (memset / fill), with torch's own kernels on
2 GiB buffers -- the context for kernels that mostly write (pair_hidden_fwd, pair_features_dropout).  Prints one JSON
line.  Measured on the pool's B200: copy 6578 GB/s, write only 3934-3943 GB/s (60 % of the copy figure)."""
import json
import torch

n = 1 << 30   # fp16 elements: 2 GiB
a = torch.empty(n, device='cuda', dtype=torch.float16)
b = torch.empty(n, device='cuda', dtype=torch.float16)
a.fill_(1.0)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


gb = 2.0 * n / 1e9
out = {'copy_gbs (read + write bytes)': 2 * gb / timeit(lambda: b.copy_(a)),
       'write_only_gbs (memset)': gb / timeit(lambda: b.zero_()),
       'write_only_gbs (fill)': gb / timeit(lambda: b.fill_(2.0))}
print(json.dumps({k: round(v, 1) for k, v in out.items()}))
