cd /root/repo
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/r1d_bench_default.json 2> gpurun_out/r1d_bench_default.err
python tools/show_bench.py gpurun_out/r1d_bench_default.json | head -12
timeout 300 python bench.py --no-cpu-baseline --workload c3 > gpurun_out/r1d_bench_train_c3.json 2>/dev/null
python tools/show_bench.py gpurun_out/r1d_bench_train_c3.json | head -1
