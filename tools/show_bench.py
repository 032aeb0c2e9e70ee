"""Prints the headline numbers and the per-kernel table of a bench.py JSON line."""
import json
import sys

d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print('value %.0f q/s  %.3f ms/step | e2e %.0f q/s %.3f ms | launches %d' % (
    d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches']))
r = d['roofline']
print('roofline: %s %s %.1f/%.1f %s frac %.3f share %.3f' % (r['kernel'], r['bound'], r['achieved'], r['peak'],
                                                            r['unit'], r['frac'], r['share_of_kernel_time']))
for k, v in d['kernels'].items():
    print('  %-44s %8.4f ms x%-4.1f %s %s' % (k, v['ms_per_step'], v['launches_per_step'],
                                          ('%6.0f TF/s' % v['tflops']) if v['tflops'] else '',
                                          ('%6.0f GB/s' % v['gbs']) if v['gbs'] else ''))
