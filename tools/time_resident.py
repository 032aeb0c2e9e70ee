"""Times the persistent pair-layer GEMMs alone at c1 size (P = 589824), optionally under DFOL_RS_DEBUG."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfol_vqa_b200.capi import call, ptr, stream_ptr

P, K, E = 589824, 256, 300
g = torch.Generator().manual_seed(0)
A = (torch.randn(P, K, generator=g) * 0.5).cuda().bfloat16()
W2 = (torch.randn(E, K, generator=g) / 16).cuda().bfloat16()
b2 = torch.randn(E, generator=g).cuda()
H2 = torch.empty(P, 320, device='cuda', dtype=torch.bfloat16)
dZ = torch.randn(P, 320, device='cuda').bfloat16()
Wt = (torch.randn(256, 320, generator=g) / 16).cuda().bfloat16()
dX = torch.empty(P, 256, device='cuda', dtype=torch.bfloat16)
flush = torch.empty(256 << 20, device='cuda', dtype=torch.uint8)


def timeit(fn, n=5):
    ts = []
    for i in range(n + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts[2:])


def fwd(store=True):
    call('dfol_pair_layer_fwd_tc', ptr(A), K, ptr(W2), K, ptr(H2) if store else None, 320, 320, ptr(b2), P, E, K, 2,
         None, 0, None, None, None, 0, None, None, None, None, None, 0.0, None, stream_ptr())


def dgrad():
    call('dfol_pair_layer_dgrad_tc', ptr(dZ), 320, ptr(Wt), 320, ptr(dX), 256, 0, P, 256, 320, ptr(A), 256, 2, stream_ptr())


def fwd_cluster():
    call('dfol_pair_layer_fwd_cluster', ptr(A), K, ptr(W2), K, ptr(H2), 320, 320, ptr(b2), P, E, K, 2, stream_ptr())


def dgrad_cluster():
    call('dfol_pair_layer_dgrad_cluster', ptr(dZ), 320, ptr(Wt), 320, ptr(dX), 256, 0, P, 256, 320, ptr(A), 256, 2, 1.0,
         stream_ptr())


print('cluster: fwd %.3f ms  dgrad %.3f ms' % (timeit(fwd_cluster), timeit(dgrad_cluster)))
print('(DFOL_RS_DEBUG needs a -DDFOL_RS_EXPERIMENTS build) DFOL_RS_DEBUG=%s fwd %.3f ms  dgrad %.3f ms' % (
    os.environ.get('DFOL_RS_DEBUG', '0'), timeit(fwd), timeit(dgrad)))
