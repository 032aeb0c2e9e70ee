cd /root/repo
for k in 1 8 16 32; do
DFOL_WG_MIN_KB=$k timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_wg$k.json 2>gpurun_out/wg.err
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_wg$k.json') if x.startswith('{')][-1]
d=json.loads(l)
print('min_kb=$k', round(d['ms_per_step'],4), ' '.join('%s=%.4f' % (k.replace('gemm_bf16_tc_',''), v['ms_per_step']) for k,v in d['kernels'].items() if 'wgrad' in k))
PY
done
