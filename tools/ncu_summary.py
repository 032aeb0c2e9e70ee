"""Summarises `ncu --page raw --csv` exports: a markdown table per capture and profiles/r2_traffic.json
(DRAM bytes per launch of the kernels bench.py names, keyed ``<bench tag>@<workload>``)."""
import csv
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'us': 1e-3, 'ms': 1.0, 'ns': 1e-6, 's': 1e3}
TAGS = [('gemm_bf16_tc_cluster_kernel<2, 0>', 'pair_layer_fwd_cluster[Px300x256]'),
        ('gemm_bf16_tc_cluster_kernel<2, 0, 0>', 'pair_layer_fwd_cluster[Px300x256]'),
        ('gemm_bf16_tc_cluster_kernel<0, 1>', 'pair_layer_dgrad_cluster[Px256x320]'),
        ('gemm_bf16_tc_cluster_kernel<0, 1, 0>', 'pair_layer_dgrad_cluster[Px256x320]'),
        ('gemm_bf16_tc_cluster_kernel<0, 1, 1>', 'pair_layer_dgrad_wgrad_cluster[Px256x320]'),
        ('table_layer_bwd_mma_kernel', 'table_layer_bwd_mma[rel]'),
        ('pair_hidden_fwd_tc_kernel', 'pair_hidden_fwd_tc'), ('pair_hidden_bwd', 'pair_hidden_bwd_tc'),
        ('rel_slots_tc_kernel', 'rel_slots_fwd_tc'), ('program_fwd_kernel', 'program_fwd'),
        ('program_bwd_kernel', 'program_bwd'), ('pair_chain_fwd_kernel', 'pair_chain_fwd[Px300x256]')]


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    return rows[hi], rows[hi + 1], rows[hi + 2:]


def main():
    traffic = {}
    out = []
    for workload in sys.argv[1:]:
        path = os.path.join(REPO, 'profiles', 'r2_prof_train_bf16_%s.raw.csv' % workload)
        hdr, units, rows = load(path)
        ix = {h: i for i, h in enumerate(hdr)}

        def val(r, key):
            return float(r[ix[key]].replace(',', '')) * SCALE.get(units[ix[key]], 1.0)

        out.append('\n### %s (`%s`)\n' % (workload, os.path.basename(path)))
        out.append('| kernel | grid | time ms | dram read MB | dram write MB | dram % | tensor pipe % | SM % | regs |')
        out.append('|---|---|---|---|---|---|---|---|---|')
        seen = {}
        for r in rows:
            name = r[ix['Kernel Name']]
            short = name.split('(')[0].replace('void ', '')
            t = val(r, 'gpu__time_duration.sum')
            rd, wr = val(r, 'dram__bytes_read.sum'), val(r, 'dram__bytes_write.sum')
            out.append('| %s | %s | %.3f | %.1f | %.1f | %.1f | %.1f | %.1f | %s |' % (
                short[:60], r[ix['launch__grid_size']], t, rd / 1e6, wr / 1e6,
                float(r[ix['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']]),
                float(r[ix['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']]),
                float(r[ix['sm__throughput.avg.pct_of_peak_sustained_elapsed']]), r[ix['launch__registers_per_thread']]))
            for pat, tag in TAGS:
                if pat in name:
                    key = '%s@%s' % (tag, workload)
                    # the largest launch of the kind (the pair-level one) is the kernel bench.py names
                    if rd + wr > seen.get(key, 0):
                        seen[key] = rd + wr
                        traffic[key] = rd + wr
            if 'gemm_bf16_tc_wgrad_kernel' in name and rd + wr > max(2e8, traffic.get('gemm_bf16_tc_wgrad[300x256xP]@' + workload, 0)):   # (pair-level launches only)
                traffic['gemm_bf16_tc_wgrad[300x256xP]@' + workload] = rd + wr
            if 'table_layer_bwd_tc_kernel' in name and rd + wr > max(1e8, traffic.get('table_layer_bwd_tc[rel]@' + workload, 0)):
                traffic['table_layer_bwd_tc[rel]@' + workload] = rd + wr
    json.dump(traffic, open(os.path.join(REPO, 'profiles', 'r2_traffic.json'), 'w'), indent=1, sort_keys=True)
    print('\n'.join(out))


if __name__ == '__main__':
    main()
