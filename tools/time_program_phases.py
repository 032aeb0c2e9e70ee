"""Timing experiment: per-phase clock64 stamps of one question block of the fast forward interpreter.
Needs a library whose program_fwd_fast.cu was compiled with -DDFOL_PROG_TIMING (tools/run_phase_timing.sh builds it):
  python tools/time_program_phases.py /tmp/libdfol_timing.so [workload] [batch]"""
import ctypes, os, sys
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tools'))
from dfol_vqa_b200 import capi
capi.LIB_PATH = sys.argv[1]
from prog_setup import setup
interp, eng, cp, scene = setup(sys.argv[2] if len(sys.argv) > 2 else 'c3', int(sys.argv[3]) if len(sys.argv) > 3 else 0)
with torch.no_grad():
    for _ in range(3):
        eng.run_programs(cp, scene, save_tape=True)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 512)()
h = ctypes.CDLL(sys.argv[1])
h.dfol_prog_timing_read(buf, 512)
t = np.array(buf[:9 * 8]).reshape(9, 8)
names = ['A:prior+exp', 'wait tile', 'barrier', 'hop(products+finish)', 'refill']
for k in range(9):
    d = np.diff(t[k, :6])
    print('hop %d: ' % k + '  '.join('%s %5d' % (nm, v) for nm, v in zip(names, d)) + '  | total %6d  | gap to next %s' % (
        t[k, 5] - t[k, 0], (t[k + 1, 0] - t[k, 5]) if k < 8 else '-'))
print('entry -> loop start %d cycles; loop start -> first relate %d; last relate end -> kernel end %d; entry -> end %d' % (t[0,7]-t[0,6], t[0,0]-t[0,7], t[1,6]-t[8,5], t[1,6]-t[0,6]))
