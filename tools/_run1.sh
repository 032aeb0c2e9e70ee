cd /root/repo
timeout 900 python -m pytest tests/test_gpu_tc_kernels.py tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | tail -2
for w in c1 c3; do
for v in new old; do
if [ $v = old ]; then export DFOL_LIB_PATH=/root/repo/tools/_old/libdfol_b200.so; else unset DFOL_LIB_PATH; fi
timeout 300 python bench.py --no-cpu-baseline --workload $w > gpurun_out/bench_sb_${v}_$w.json 2>gpurun_out/sb.err
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_sb_${v}_$w.json') if x.startswith('{')][-1]
d=json.loads(l)
print('$v $w', round(d['ms_per_step'],4), ' '.join('%s=%.4f' % (k.replace('table_layer_bwd_tc','tbl'), v['ms_per_step']) for k,v in d['kernels'].items() if 'table_layer' in k))
PY
done
done
