"""Per-parameter gradient error of both gemm modes against the CPU oracle on a bench.py workload sub-sample."""
import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import torch
import helpers
import dfol_oracle as orc
import bench
from dfol_vqa_b200 import synth
from dfol_vqa_b200.ontology import synthetic_ontology
from dfol_vqa_b200.programs import ProgramCollater
from dfol_vqa_b200.interpreter import FusedTrainStep

workload, seed, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 32
wl = bench.WORKLOADS[workload]
ont = synthetic_ontology(seed=1, embedding_dim=300, **bench.VOCAB)
questions = bench.make_workload_questions(ont, wl, B, seed, index=seed)
feats, bidx = synth.make_object_features([wl['n']] * B, 2048, seed=seed + 7)
coll = lambda: ProgramCollater(1, lambda qs: (feats, bidx)).collate(json.loads(json.dumps(questions)))
for mode in ('fp32', 'bf16'):
    interp = helpers.build_interpreter(ont, bench.DIMS, seed=0, gemm_mode=mode, emb_bias=bench.EMB_BIAS)
    params = helpers.oracle_params(interp, torch.float32, requires_grad=True)
    res, loss_ref = orc.run_step(ont, params, coll(), is_training=True)
    loss_ref.backward()
    step = FusedTrainStep(interp)
    pbs = helpers.to_cuda(coll())
    loss = step.forward_backward(pbs)
    interp.eval()
    with torch.no_grad():
        lp = interp(pbs, False)['log_probability'].cpu()
    lp_ref = res[0]['log_probability'].detach()
    print(mode, 'loss', float(loss), float(loss_ref), 'lp max err', float((lp - lp_ref).abs().max()))
    if mode == 'fp32':
        print('lp_ref', [round(float(v), 4) for v in lp_ref[:40]])
    keys = {id(p): k for k, p in interp.named_parameters()}
    for p in interp.oracle_parameters():
        k = keys[id(p)]
        g_ref = params[k].grad if params[k].grad is not None else torch.zeros_like(params[k])
        got = step.grads[id(p)].cpu()
        e = (got - g_ref).abs()
        print('   %-55s scale %.3e maxerr %.3e (%.1f%%) fro %.4f' % (k, float(g_ref.abs().max()), float(e.max()),
              100 * float(e.max()) / max(float(g_ref.abs().max()), 1e-30),
              float((got - g_ref).norm()) / max(float(g_ref.norm()), 1e-30)))
