"""Times the interpreter forward kernel alone on a built c3 / c1 scene: L2-cold (flushed) vs L2-warm."""
import os
import sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, 'tests'))
import argparse
import bench

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='c3')
ap.add_argument('--batch', type=int, default=0)
a = ap.parse_args()
args = argparse.Namespace(workload=a.workload, gemm='bf16', local_batch=a.batch, pool=1, mode='infer')
dev = torch.device('cuda', 0)
ont, interp, host_batches, B = bench.build_world(args, 0, dev)
pb = host_batches[0].to_cuda(0)
from dfol_vqa_b200.engine import SceneLayout
cp = interp.compiled(pb, False)
counts = interp._object_counts(pb)
layout = SceneLayout.get(counts, interp._weights.emb.weight.shape[0], len(ont._relation_index), dev)
eng = interp._engine
with torch.no_grad():
    scene = eng.build_scene(pb._object_features.float(), layout, keep_for_backward=False, cp=cp)
flush = torch.empty(512 << 20, device='cuda', dtype=torch.uint8)


def timeit(cold, n=6):
    ts = []
    for i in range(n):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.run_programs(cp, scene, save_tape=False)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts[1:])


cold, warm = timeit(True), timeit(False)
print('%s program_fwd: cold %.1f us (%.0f GB/s)  warm %.1f us (%.0f GB/s)  alg bytes %.1f MB' % (
    a.workload, cold * 1e3, cp.alg_bytes / cold / 1e6, warm * 1e3, cp.alg_bytes / warm / 1e6, cp.alg_bytes / 1e6))
