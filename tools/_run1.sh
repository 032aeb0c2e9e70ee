cd /root/repo
timeout 900 python -m pytest tests/test_gpu_tc_kernels.py -x -q -m gpu 2>&1 | tail -2
for k in 0 1; do
for w in c1 c3; do
DFOL_PK_TBL=$k timeout 300 python bench.py --no-cpu-baseline --workload $w > gpurun_out/bench_pk${k}_$w.json 2>gpurun_out/bench_pk_$w.err
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_pk${k}_$w.json') if x.startswith('{')][-1]
d=json.loads(l)
print('tbl pk=$k $w', round(d['ms_per_step'],4), ' '.join('%s=%.4f' % (k.replace('pair_hidden','ph').replace('table_layer_bwd_tc','tbl'), v['ms_per_step']) for k,v in d['kernels'].items() if 'pair_hidden' in k or 'table_layer' in k))
PY
done
done
