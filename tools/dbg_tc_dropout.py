import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests')); sys.path.insert(0, os.path.join(REPO, 'oracle'))
import helpers, dfol_oracle as orc
from test_gpu_tc_kernels import _programs_world
from dfol_vqa_b200.interpreter import FusedTrainStep
terminal, n_max = sys.argv[1], int(sys.argv[2])
for p in (0.1,):
    ont, dims, pbs = _programs_world(terminal, 12, n_max, True, seed=83)
    out = {}
    for mode in ('fp32', 'bf16'):
        interp = helpers.build_interpreter(ont, dims, seed=5, gemm_mode=mode, emb_bias=-4.0, dropout=p)
        interp._fixed_dropout_seed = 1717
        interp.train()
        step = FusedTrainStep(interp)
        loss = float(step.forward_backward([pbs[0].to_cuda(0)]))
        out[mode] = (loss, {k: step.grads[id(q)].clone() for k, q in zip(orc.PARAM_KEYS, interp.oracle_parameters())})
    print('dropout', p, 'loss', out['fp32'][0], out['bf16'][0])
    for k in orc.PARAM_KEYS:
        a, b = out['bf16'][1][k], out['fp32'][1][k]
        c = float((a * b).sum() / ((b * b).sum() + 1e-30))
        print('  %-52s norm-rel %.3f  proj %.3f  resid %.3f  scale %.3g' % (k, float((a - b).norm() / (b.norm() + 1e-20)), c, float((a - c * b).norm() / (b.norm() + 1e-20)), float(b.abs().max())))
