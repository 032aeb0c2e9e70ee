"""Per-tile clock stamps of the cluster GEMM roles (DFOL_CL_TRACE=1): one launch each of fwd, dgrad, fused dgrad+wgrad."""
import os
import sys
os.environ['DFOL_CL_TRACE'] = '1'
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfol_vqa_b200.capi import call, ptr, stream_ptr

P = 426 * 2304
g = torch.Generator().manual_seed(0)
H1 = (torch.rand(P, 256, device='cuda') - 0.3).bfloat16()
dZ = torch.randn(P, 320, device='cuda').bfloat16()
Wt = (torch.randn(256, 320, generator=g) / 16).cuda().bfloat16()
W2 = (torch.randn(300, 256, generator=g) / 16).cuda().bfloat16()
b2 = torch.randn(300, generator=g).cuda()
H2 = torch.empty(P, 320, device='cuda', dtype=torch.bfloat16)
dX = torch.empty(P, 256, device='cuda', dtype=torch.bfloat16)
dW = torch.zeros(300, 256, device='cuda')
st = stream_ptr()
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
if which in ('all', 'fwd'):
    call('dfol_pair_layer_fwd_cluster', ptr(H1), 256, ptr(W2), 256, ptr(H2), 320, 320, ptr(b2), P, 300, 256, 2, st)
if which in ('all', 'dgrad'):
    call('dfol_pair_layer_dgrad_cluster', ptr(dZ), 320, ptr(Wt), 320, ptr(dX), 256, 0, P, 256, 320, ptr(H1), 256, 2, 1.0, st)
if which in ('all', 'fused'):
    call('dfol_pair_layer_dgrad_wgrad_cluster', ptr(dZ), 320, ptr(Wt), 320, ptr(dX), 256, 0, P, 256, 320, ptr(H1), 256, 2,
         1.0, ptr(dW), 256, 300, st)
torch.cuda.synchronize()
