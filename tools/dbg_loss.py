import sys, os, json, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests'); sys.path.insert(0, '/root/repo/oracle')
import helpers, dfol_oracle as orc
from dfol_vqa_b200 import synth
from dfol_vqa_b200.ontology import synthetic_ontology
from dfol_vqa_b200.programs import ProgramCollater
dims = dict(box=2048, feat=512, hidden=256, emb=300)
ont = synthetic_ontology(400, 60, 6, 5, seed=3, embedding_dim=300)
for terminal in ['verify_rel', 'exist', 'and']:
    interp = helpers.build_interpreter(ont, dims, seed=5, gemm_mode='bf16', emb_bias=-4.0)
    questions = synth.make_questions(ont, 16, terminal, 1, 3, seed=31, relate_prob=0.6)
    counts = synth.object_counts(16, 48, True, seed=32)
    feats, bidx = synth.make_object_features(counts, 2048, seed=33)
    pbs = ProgramCollater(1, lambda qs: (feats, bidx)).collate(questions)
    params = helpers.oracle_params(interp, torch.float32, requires_grad=False)
    results, loss_ref = orc.run_step(ont, params, ProgramCollater(1, lambda qs: (feats, bidx)).collate(json.loads(json.dumps(questions))), is_training=True)
    interp.train()
    with torch.no_grad():
        out = interp(helpers.to_cuda(pbs), True)
    lp = out['log_probability'].cpu()
    ref = torch.cat([r['log_probability'].detach() for r in results]) if isinstance(results, list) else results['log_probability'].detach()
    print(terminal, 'loss_ref', float(loss_ref))
    for a, b, q in zip(lp.tolist(), ref.tolist(), questions):
        print('   %.6e  %.6e  ans=%s' % (a, b, q['answer']))
