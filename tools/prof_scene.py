"""Runs the c1-shape hot path a few times for ncu (python tools/prof_scene.py [infer|train] [fp32|bf16] [workload] [reps])."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import argparse

import torch

import bench

mode = sys.argv[1] if len(sys.argv) > 1 else 'infer'
gemm = sys.argv[2] if len(sys.argv) > 2 else 'bf16'
workload = sys.argv[3] if len(sys.argv) > 3 else 'c1'
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
args = argparse.Namespace(workload=workload, gemm=gemm, local_batch=0, pool=1, mode=mode)
torch.cuda.set_device(0)
ont, interp, batches, B = bench.build_world(args, 0, torch.device('cuda', 0))
pb = batches[0].to_cuda(0)
from dfol_vqa_b200.interpreter import FusedTrainStep
trainer = FusedTrainStep(interp) if mode == 'train' else None
interp.train(mode == 'train')
for _ in range(reps):
    if trainer is not None:
        trainer.step([pb])
    else:
        with torch.no_grad():
            interp([pb], True)
torch.cuda.synchronize()
print('done')
