"""Debug aid: where do the fast and exact interpreter builds disagree when modulations are on?"""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import helpers
from test_gpu_tc_kernels import _programs_world
from dfol_vqa_b200.engine import SceneLayout
from dfol_vqa_b200.networks import build_attention_networks
terminal, n_max, ragged = sys.argv[1], int(sys.argv[2]), sys.argv[3] == '1'
ont, dims, pbs = _programs_world(terminal, 12, n_max, ragged, seed=41)
nets = build_attention_networks(dims['emb'], 50)
interp = helpers.build_interpreter(ont, dims, seed=5, gemm_mode='bf16', emb_bias=-4.0,
                                   attention_nets=[nets[k] for k in ('forward_attention_network', 'backward_attention_network', 'attention_output_network')])
pb = pbs[0].to_cuda(0)
cp = interp.compiled(pb, False)
counts = interp._object_counts(pb)
layout = SceneLayout.get(counts, interp._weights.emb.weight.shape[0], len(ont._relation_index), torch.device('cuda', 0))
eng = interp._engine
torch.manual_seed(9)
rows = 0.06 + 0.08 * torch.rand(cp.mod_rows, 4, device='cuda')
rows[:, 3] = 0.3 + 0.4 * torch.rand(cp.mod_rows, device='cuda')
out = {}
with torch.no_grad():
    scene = eng.build_scene(pb._object_features.float(), layout, keep_for_backward=True, cp=cp)
    for mode in ('bf16', 'fp32'):
        eng.gemm_mode = mode
        scene.mods = rows.contiguous(); scene.d_mods = torch.zeros_like(rows)
        lp, tape = eng.run_programs(cp, scene, save_tape=True)
        d_lp = torch.linspace(-1.0, 1.0, lp.numel(), device='cuda')
        g_attr, g_rel = eng.program_backward(cp, scene, tape, d_lp)
        out[mode] = (lp.clone(), g_attr.clone(), g_rel.clone(), scene.d_mods.clone(), tape.clone())
names = ['lp', 'g_attr', 'g_rel', 'd_mods', 'tape']
for nm, a, b in zip(names, out['bf16'], out['fp32']):
    d = (a - b).abs()
    i = int(d.reshape(-1).argmax())
    print(nm, 'max|b| %.4g maxdiff %.4g at %d: fast %.6g exact %.6g' % (float(b.abs().max()), float(d.max()), i, float(a.reshape(-1)[i]), float(b.reshape(-1)[i])))
ga_f, ga_e = out['bf16'][1], out['fp32'][1]
i = int((ga_f - ga_e).abs().argmax())
for (q, col, off) in cp.attr_slices:
    if off <= i < off + layout.a_stride_host[q]:
        print('slice question', q, 'col', col, 'off', off, 'idx in slice', i - off)
        ins = cp.instr[cp.q_instr[q]:cp.q_instr[q + 1]]
        print(ins)
        sl = slice(off, off + counts[q])
        print('fast ', ga_f[sl].cpu().numpy()); print('exact', ga_e[sl].cpu().numpy())
        for k, ip in enumerate(range(cp.q_instr[q], cp.q_instr[q + 1])):
            st = out['bf16'][4].view(-1, (max(counts) + 3) // 4 * 4)
            print('tape', ip, st[ip][:counts[q]].cpu().numpy()[:12])
        break
q = 3
stf = out['bf16'][4].view(-1, (max(counts) + 3) // 4 * 4); ste = out['fp32'][4].view(-1, (max(counts) + 3) // 4 * 4)
for ip in range(cp.q_instr[q], cp.q_instr[q + 1]):
    d = (stf[ip] - ste[ip]).abs()
    print('ip', ip, 'maxdiff', float(d.max()), 'at', int(d.argmax()), 'fast', float(stf[ip][int(d.argmax())]), 'exact', float(ste[ip][int(d.argmax())]))
print('mods rows', rows[[3, 39, 27, 87, 75, 99]].cpu().numpy())
# relation tile of the second relate of question 3: slot 1
blk = int(cp.slot_blk[q]); rs = int(layout.r_stride_host[q]); n = counts[q]
tile = scene.rel_ll[blk + rs: blk + rs + n * n].view(n, n).cpu()
print('tile row 41 max', float(tile[41].max()), 'argmax', int(tile[41].argmax()), 'diag', float(tile[41, 41]))
print('cur before (tape 13) fast', stf[13][:n].cpu().numpy()[[int(tile[41].argmax()), 41]])
