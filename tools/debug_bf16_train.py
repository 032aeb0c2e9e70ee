import json, os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests')); sys.path.insert(0, os.path.join(REPO, 'oracle'))
import torch
import helpers
from dfol_vqa_b200 import synth, capi
from dfol_vqa_b200.ontology import synthetic_ontology
from dfol_vqa_b200.programs import ProgramCollater
from dfol_vqa_b200.interpreter import FusedTrainStep
from dfol_vqa_b200.engine import SceneLayout
dims = dict(box=2048, feat=512, hidden=256, emb=300)
ont = synthetic_ontology(400, 60, 6, 5, seed=3, embedding_dim=300)
interp = helpers.build_interpreter(ont, dims, seed=5, gemm_mode='bf16', emb_bias=-4.0)
questions = synth.make_questions(ont, 16, 'verify_rel', 1, 3, seed=31, relate_prob=0.6)
counts = synth.object_counts(16, 48, True, seed=32)
feats, bidx = synth.make_object_features(counts, 2048, seed=33)
pb = ProgramCollater(1, lambda qs: (feats, bidx)).collate(questions)[0].to_cuda(0)
eng = interp._engine
lay = SceneLayout.get(counts, 400, 60, torch.device('cuda', 0))
cp = interp.compiled(pb, False)
scene = eng.build_scene(pb._object_features, lay)
def chk(name, t): print('%-12s nan=%d inf=%d absmax=%.3e' % (name, int(torch.isnan(t.float()).sum()), int(torch.isinf(t.float()).sum()), float(t.float().abs().max())))
for n in ['obj', 'uv', 'attr_ll', 'rel_ll']: chk(n, getattr(scene, n))
for i, h in enumerate(scene.rel_h): chk('rel_h%d' % i, h)
lp, tape = eng.run_programs(cp, scene); chk('lp', lp)
# monkeypatch call to check outputs after each backward kernel
step = FusedTrainStep(interp)
orig = capi.call
import dfol_vqa_b200.engine as E
import ctypes
def traced(name, *args):
    orig(name, *args); torch.cuda.synchronize()
    bad = [n for n, t in list(step.grads.items()) if not torch.isfinite(t).all()]
    print('  called', name, 'nonfinite grads:', len(bad))
E.call = traced
from dfol_vqa_b200.interpreter import targets_of
target = torch.from_numpy(targets_of(cp, pb._answers)).cuda()
d_lp = torch.empty_like(lp); scal = torch.zeros(2, device='cuda')
capi.call('dfol_loss_fwd_bwd', capi.ptr(lp), capi.ptr(target), None, cp.question_num, cp.lp_num, cp.kind, 1.0/16, capi.ptr(scal), capi.ptr(d_lp), capi.stream_ptr())
print('lp', lp.tolist()); print('target', target.tolist()); print('d_lp', d_lp.tolist()); print('loss', scal.tolist())
step.flat_grad.zero_()
# wrap helper tensors: run backward and check grads
eng.backward(cp, scene, tape, d_lp, step.grads)
keys = {id(p): k for k, p in interp.named_parameters()}
for p in interp.oracle_parameters(): chk(keys[id(p)][-40:], step.grads[id(p)])
