import os, sys, torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
from test_gpu_tc_kernels import _programs_world
from dfol_vqa_b200.compiler import ProgramCompiler
from dfol_vqa_b200.modulator import AttentionTransfer
from dfol_vqa_b200.modulator_cuda import NativeAttentionTransfer
from dfol_vqa_b200.networks import build_attention_networks
terminal, n_max, S = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ont, dims, pbs = _programs_world(terminal, 24, n_max, True, seed=77)
torch.manual_seed(11)
nets = build_attention_networks(dims['emb'], S)
with torch.no_grad():
    nets['attention_output_network'][0].weight.normal_(0.0, 0.3)
fwd, bwd, out = (nets[k].cuda() for k in ('forward_attention_network', 'backward_attention_network', 'attention_output_network'))
cp = ProgramCompiler(ont, normalize=True, modulated=True).compile(pbs[0], [n_max] * 24)
ref = AttentionTransfer(fwd, bwd, out, ont)
rows = ref.modulations(cp).detach()
native = NativeAttentionTransfer(fwd, bwd, out, ont)
mods, ctx = native.forward(cp)
for (slot, key, n, base) in cp.mod_plan:
    d = cp.mod_descs[slot]
    diff = (mods[base:base+n] - rows[base:base+n]).abs().max(dim=1)[0]
    bad = (diff > 1e-5).nonzero().flatten().tolist()
    print(slot, d['op'], key, n, 'deps', d['deps'], 'maxdiff %.3g' % float(diff.max()), 'bad rows', bad[:12], 'mask', None if d['mask'] is None else [int(m) for m in d['mask']])
