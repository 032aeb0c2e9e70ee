cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline --workload c2 > gpurun_out/r1d_bench_train_c2.json 2>gpurun_out/c2.err; tail -3 gpurun_out/c2.err
python - <<PY
import json
l=[x for x in open('gpurun_out/r1d_bench_train_c2.json') if x.startswith('{')][-1]
d=json.loads(l)
print('c2', round(d['ms_per_step'],4), round(d['value']), 'e2e', round(d['e2e']['value']))
for k,v in list(d['kernels'].items())[:14]: print('  ',k, round(v['ms_per_step'],4))
PY
