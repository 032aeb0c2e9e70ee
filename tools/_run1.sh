cd /root/repo
echo "== old path"; DFOL_DENSE_FP32=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "bf16_mode_training_gradients" 2>&1 | grep -i "assert\|passed\|failed" | head
echo "== new path"; timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "bf16_mode_training_gradients" 2>&1 | grep -i "assert\|passed\|failed" | head
timeout 300 python bench.py --no-cpu-baseline --workload c2 > gpurun_out/r1d_bench_train_c2.json 2>gpurun_out/c2.err; tail -3 gpurun_out/c2.err
python - <<PY
import json
l=[x for x in open('gpurun_out/r1d_bench_train_c2.json') if x.startswith('{')][-1]
d=json.loads(l)
print('c2', round(d['ms_per_step'],4), round(d['value']), 'e2e', round(d['e2e']['value']))
for k,v in list(d['kernels'].items())[:8]: print('  ',k, round(v['ms_per_step'],4))
PY
