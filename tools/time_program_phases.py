"""Timing experiment: per-phase clock64 stamps of one question block of the fast forward interpreter.
Needs a library whose program_fwd_fast.cu was compiled with -DDFOL_PROG_TIMING:
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -Iinclude -Idfol_vqa_b200/csrc \
       -DDFOL_PROG_TIMING -c dfol_vqa_b200/csrc/program_fwd_fast.cu -o /tmp/pf.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o tools/micro/libdfol_timing.so \
       $(ls dfol_vqa_b200/build/*.o | grep -v program_fwd_fast.o) /tmp/pf.o -lcudart
  python tools/time_program_phases.py tools/micro/libdfol_timing.so [batch]"""
import ctypes, os, sys
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
from dfol_vqa_b200 import capi
capi.LIB_PATH = sys.argv[1]
import argparse, bench
args = argparse.Namespace(workload='c3', gemm='bf16', local_batch=int(sys.argv[2]) if len(sys.argv) > 2 else 0, pool=1, mode='infer')
dev = torch.device('cuda', 0)
ont, interp, host_batches, B = bench.build_world(args, 0, dev)
pb = host_batches[0].to_cuda(0)
from dfol_vqa_b200.engine import SceneLayout
cp = interp.compiled(pb, False)
layout = SceneLayout.get(interp._object_counts(pb), interp._weights.emb.weight.shape[0], len(ont._relation_index), dev)
eng = interp._engine
with torch.no_grad():
    scene = eng.build_scene(pb._object_features.float(), layout, keep_for_backward=False, cp=cp)
    for _ in range(3):
        eng.run_programs(cp, scene, save_tape=False)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 512)()
h = ctypes.CDLL(sys.argv[1])
h.dfol_prog_timing_read(buf, 512)
t = np.array(buf[:9 * 8]).reshape(9, 8)
names = ['A:prior+exp', 'wait tile', 'barrier', 'products', 'C:finish+barrier']
for k in range(9):
    d = np.diff(t[k, :6])
    print('hop %d: ' % k + '  '.join('%s %5d' % (nm, v) for nm, v in zip(names, d)) + '  | total %6d  | gap to next %s' % (
        t[k, 5] - t[k, 0], (t[k + 1, 0] - t[k, 5]) if k < 8 else '-'))
print('entry -> loop start %d cycles; loop start -> first relate %d; last relate end -> kernel end %d; entry -> end %d' % (t[0,7]-t[0,6], t[0,0]-t[0,7], t[1,6]-t[8,5], t[1,6]-t[0,6]))
