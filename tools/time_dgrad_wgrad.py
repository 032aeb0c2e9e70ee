"""Times the pair-level second-layer backward alone: cluster dgrad, stand-alone wgrad, and the fused dgrad + wgrad, at the
pair-row counts of c1 / c4 / c3 (inputs larger than L2, L2 flushed between launches)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfol_vqa_b200.capi import call, ptr, stream_ptr

flush = torch.empty(256 << 20, device='cuda', dtype=torch.uint8)


def timeit(fn, n=6):
    ts = []
    for i in range(n + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:])
    return ts[len(ts) // 2]


for name, P in (('c1', 169 * 2304), ('c4', 426 * 2304), ('c3', 256 * 10000)):
    g = torch.Generator().manual_seed(0)
    H1 = (torch.rand(P, 256, device='cuda') - 0.3).bfloat16()
    dZ = torch.randn(P, 320, device='cuda').bfloat16()
    dZ[:, 300:] = 0
    Wt = (torch.randn(256, 320, generator=g) / 16).cuda().bfloat16()
    dX = torch.empty(P, 256, device='cuda', dtype=torch.bfloat16)
    dW = torch.zeros(300, 256, device='cuda')
    st = stream_ptr()
    t_d = timeit(lambda: call('dfol_pair_layer_dgrad_cluster', ptr(dZ), 320, ptr(Wt), 320, ptr(dX), 256, 0, P, 256, 320,
                              ptr(H1), 256, 2, 1.0, st))
    t_w = timeit(lambda: call('dfol_gemm_bf16_tc_wgrad', ptr(dZ), 320, ptr(H1), 256, ptr(dW), 256, 300, 256, P, st))
    t_f = timeit(lambda: call('dfol_pair_layer_dgrad_wgrad_cluster', ptr(dZ), 320, ptr(Wt), 320, ptr(dX), 256, 0, P, 256,
                              320, ptr(H1), 256, 2, 1.0, ptr(dW), 256, 300, st))
    gb = 2.0 * P * (320 + 512) / 1e9
    print('%s P=%d: dgrad %.3f ms + wgrad %.3f ms = %.3f | fused %.3f ms (%.0f GB/s on the dgrad bytes) env EC=%s PF=%s' % (
        name, P, t_d, t_w, t_d + t_w, t_f, gb / t_f * 1e3, os.environ.get('DFOL_CL_EC', '-'),
        os.environ.get('DFOL_CL_PREFETCH', '-')))
    del H1, dZ, dX
