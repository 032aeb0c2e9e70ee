"""Runs the interpreter kernels of one bench.py workload in a loop (target of ncu captures / quick timing):
  python tools/prog_loop.py workload [reps] [batch]"""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
from prog_setup import setup
interp, eng, cp, scene = setup(sys.argv[1], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
with torch.no_grad():
    lp, tape = eng.run_programs(cp, scene, save_tape=True)
    d_lp = torch.full_like(lp, 1e-3)
    for _ in range(2):
        eng.run_programs(cp, scene, save_tape=True); eng.program_backward(cp, scene, tape, d_lp)
    torch.cuda.synchronize()
    ef = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for _ in range(reps):
        ef[0].record(); eng.run_programs(cp, scene, save_tape=True); ef[1].record()
        eng.program_backward(cp, scene, tape, d_lp); ef[2].record()
        torch.cuda.synchronize()
        tf += ef[0].elapsed_time(ef[1]); tb += ef[1].elapsed_time(ef[2])
print('%s: program_fwd %.1f us, program_bwd %.1f us (incl. launch + output allocation), alg bytes %.1f MB' % (
    sys.argv[1], 1e3 * tf / reps, 1e3 * tb / reps, cp.alg_bytes / 1e6))
