import os, sys, json, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests')); sys.path.insert(0, os.path.join(REPO, 'oracle'))
import helpers, dfol_oracle as orc
from test_gpu_tc_kernels import _programs_world
from test_gpu_dropout import oracle_masks
from dfol_vqa_b200.interpreter import FusedTrainStep
terminal, n_max = sys.argv[1], int(sys.argv[2])
p, seed = 0.1, 1717
ont, dims, pbs = _programs_world(terminal, 12, n_max, True, seed=83)
out = {}
for mode in ('fp32', 'bf16'):
    interp = helpers.build_interpreter(ont, dims, seed=5, gemm_mode=mode, emb_bias=-4.0, dropout=p)
    interp._fixed_dropout_seed = seed
    interp.train()
    step = FusedTrainStep(interp)
    dev_pb = pbs[0].to_cuda(0)
    loss = float(step.forward_backward([dev_pb]))
    with torch.no_grad():
        lp = interp([dev_pb], True)['log_probability'].cpu()
    out[mode] = (loss, {k: step.grads[id(q)].cpu().clone() for k, q in zip(orc.PARAM_KEYS, interp.oracle_parameters())}, lp)
counts = interp._object_counts(dev_pb)
masks = oracle_masks(interp, counts, seed, p)
params = helpers.oracle_params(interp, torch.float32, requires_grad=True)
ref = orc.OracleInterpreter(ont, params).run(pbs[0], True, masks=masks)
answers = pbs[0]._answers
loss = orc.compute_loss([ref], [answers]) / len(answers)
loss.backward()
print('loss oracle %.5f fp32 %.5f bf16 %.5f' % (float(loss), out['fp32'][0], out['bf16'][0]))
print('lp oracle', ref['log_probability'].detach().numpy().round(3))
print('lp fp32  ', out['fp32'][2].numpy().round(3))
print('lp bf16  ', out['bf16'][2].numpy().round(3))
print('answers', answers)
for k in orc.PARAM_KEYS:
    g = params[k].grad
    for mode in ('fp32', 'bf16'):
        a = out[mode][1][k]
        print('  %-50s %s norm-rel vs oracle %.4f' % (k[-40:], mode, float((a - g).norm() / (g.norm() + 1e-20))))
