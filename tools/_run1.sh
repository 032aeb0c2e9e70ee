cd /root/repo
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for w in c1 c3; do
timeout 300 python bench.py --no-cpu-baseline --workload $w > gpurun_out/bench_r1d_$w.json 2>gpurun_out/bench_r1d_$w.err
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_r1d_$w.json') if x.startswith('{')][-1]
d=json.loads(l)
print('$w', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'])
for k,v in d['kernels'].items():
    if 'pair_hidden' in k or 'table_layer' in k: print('  ',k,v['ms_per_step'],v['gbs'])
PY
done
