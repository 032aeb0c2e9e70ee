"""Wall-clock split of the calibrator training step (sample_config arrangement): modulator network forward / backward
(torch ops) vs the CUDA path."""
import os, sys, time
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench
class A: pass
args = A(); args.workload = 'c1'; args.gemm = 'bf16'; args.local_batch = 0; args.pool = 3; args.calibrate = True; args.dropout = 0.1
torch.cuda.set_device(0)
ont, interp, batches, B = bench.build_world(args, 0, torch.device('cuda', 0))
from dfol_vqa_b200.interpreter import FusedTrainStep
pbs = [pb.to_cuda(0) for pb in batches]
step = FusedTrainStep(interp)
interp.train()
for pb in pbs: step.step([pb])
torch.cuda.synchronize()
def t(fn, n=10):
    torch.cuda.synchronize(); t0 = time.time()
    for i in range(n): fn(i)
    torch.cuda.synchronize(); return (time.time() - t0) / n * 1e3
cps = [interp.compiled(pb, False) for pb in pbs]
print('slots', [len(cp.mod_descs) for cp in cps], 'rows', [cp.mod_rows for cp in cps])
print('full step            %.2f ms' % t(lambda i: step.step([pbs[i % 3]])))
def fwd_only(i):
    interp._attention.forward(cps[i % 3])
print('native modulator fwd      %.2f ms' % t(fwd_only))
grads = {id(p): torch.zeros_like(p) for p in interp.attention_parameters()}
def fwd_bwd(i):
    m, c = interp._attention.forward(cps[i % 3]); interp._attention.backward(c, torch.ones_like(m), grads)
print('native modulator fwd+bwd  %.2f ms' % t(fwd_bwd))
tm = interp._attention._torch
def torch_fwd_bwd(i):
    m = tm.modulations(cps[i % 3]); m.backward(torch.ones_like(m))
print('torch modulator fwd+bwd   %.2f ms' % t(torch_fwd_bwd))
interp._attention = None; interp._has_modulator = False
step2 = step
def no_mod(i):
    step.forward_backward([pbs[i % 3]])
try:
    print('step without modulator (fwd+program bwd) %.2f ms' % t(no_mod))
except Exception as e:
    print('no-mod failed', e)
