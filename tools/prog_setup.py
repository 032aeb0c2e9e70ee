"""Shared setup of the interpreter-kernel tools: one device batch of a bench.py workload, its scene and compiled programs."""
import argparse
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def setup(workload, batch=0, index=0, training=True):
    args = argparse.Namespace(gemm='bf16', calibrate=False, train_dropout=0.0, dropout=0.0)
    dev = torch.device('cuda', 0)
    ont, interp = bench.build_model(args, dev)
    wl = bench.WORKLOADS[workload]
    B = batch or wl['batch']
    from dfol_vqa_b200.programs import attach_compiled
    from dfol_vqa_b200.engine import SceneLayout
    host, _ = bench.build_batches(ont, wl, B, 0, index + 1)
    pb = host[index]
    attach_compiled(pb, interp._compiler, give_answer=not training)
    pb = pb.to_cuda(0)
    cp = interp.compiled(pb, not training)
    layout = SceneLayout.of_compiled(cp, dev)
    eng = interp._engine
    with torch.no_grad():
        scene = eng.build_scene(pb._object_features.float(), layout, keep_for_backward=training, cp=cp)
    torch.cuda.synchronize()
    return interp, eng, cp, scene
