"""Debug aid: fast vs exact interpreter on one synthetic case, per-instruction attention (tape) differences.
  python tools/debug_fast_vs_exact.py terminal n_max ragged neg_rel ptab"""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import helpers
from test_gpu_tc_kernels import _programs_world
from dfol_vqa_b200.engine import SceneLayout

terminal, n_max, ragged, neg_rel, ptab = sys.argv[1], int(sys.argv[2]), sys.argv[3] == '1', sys.argv[4] == '1', sys.argv[5] == '1'
ont, dims, pbs = _programs_world(terminal, 12, n_max, ragged, seed=41, neg_rel=neg_rel)
interp = helpers.build_interpreter(ont, dims, seed=5, gemm_mode='bf16', emb_bias=-4.0)
pb = pbs[0].to_cuda(0)
cp = interp.compiled(pb, False)
counts = interp._object_counts(pb)
layout = SceneLayout.get(counts, interp._weights.emb.weight.shape[0], len(ont._relation_index), torch.device('cuda', 0))
eng = interp._engine
with torch.no_grad():
    scene = eng.build_scene(pb._object_features.float(), layout, keep_for_backward=True, cp=cp)
    if not ptab:
        scene.rel_p = None
    out = {}
    for mode in ('bf16', 'fp32'):
        eng.gemm_mode = mode
        lp, tape = eng.run_programs(cp, scene, save_tape=True)
        out[mode] = (lp.clone().cpu(), tape.clone().cpu())
stride = (max(counts) + 3) // 4 * 4
lf, tf = out['bf16']; le, te = out['fp32']
print('lp fast ', lf.numpy().round(4)); print('lp exact', le.numpy().round(4))
q_instr = cp.q_instr
for q in range(cp.question_num):
    n = counts[q]
    for ip in range(int(q_instr[q]), int(q_instr[q + 1])):
        a = tf[ip * stride: ip * stride + n]; b = te[ip * stride: ip * stride + n]
        d = float((a - b).abs().max()); dp = float((a.exp() - b.exp()).abs().max())
        w = cp.instr[ip]
        flag = ' <<<<' if d > 1e-2 * (1 + float(b.abs().max())) and dp > 1e-6 else ''
        print('q%d n=%d ip=%d op=%d flags=%d a0=%d a1=%d  max|d|=%.3e max|dp|=%.3e%s' % (q, n, ip, w[0], w[1], w[2], w[3], d, dp, flag))
        if flag:
            j = int((a - b).abs().argmax())
            print('    fast', a[max(0, j - 3): j + 4].numpy(), '\n    exact', b[max(0, j - 3): j + 4].numpy())
