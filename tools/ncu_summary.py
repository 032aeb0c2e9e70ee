"""Summarises ncu outputs into profiles/: python tools/ncu_summary.py raw <rep> | launches <csv>"""
import collections
import csv
import subprocess
import sys

METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
           'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
           'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'sm__cycles_elapsed.avg',
           'smsp__inst_executed.sum', 'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum', 'lts__t_bytes.sum']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print('| kernel | grid | ' + ' | '.join(m.split('.')[0].replace('__', ':') for m in METRICS) + ' |')
    print('|---|---|' + '---|' * len(METRICS))
    for r in rows[2:]:
        name = r[idx['Kernel Name']].split('(')[0][-60:]
        vals = []
        for m in METRICS:
            vals.append('%s %s' % (r[idx[m]], units[idx[m]]) if m in idx else '-')
        print('| %s | %s | %s |' % (name, r[idx['Grid Size']], ' | '.join(vals)))


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and r[0].isdigit()]
    hdr = None
    for r in csv.reader(open(path)):
        if r and r[0] == 'ID':
            hdr = r
            break
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    total = 0.0
    for r in rows:
        name = r[idx['Kernel Name']].split('(')[0]
        val = float(r[idx['Metric Value']].replace(',', ''))
        unit = r[idx['Metric Unit']]
        scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(unit, 1e-6)
        ms = val * scale
        d = agg.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += ms
        total += ms
    print('| kernel | launches | total ms | share |')
    print('|---|---|---|---|')
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('| %s | %d | %.3f | %.1f%% |' % (name[-70:], n, ms, 100 * ms / total))
    print('| TOTAL | %d | %.3f | 100%% |' % (len(rows), total))


if __name__ == '__main__':
    {'raw': raw, 'launches': launches}[sys.argv[1]](sys.argv[2])
