#!/usr/bin/env python
"""Print the headline fields and the per-pass table of a bench.py JSON line."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
for k in ('value', 'ms_per_step', 'clocks', 'e2e', 'gpu_launches', 'roofline', 'pipeline'):
    print(k, d.get(k))
for k, v in d.get('passes', {}).items():
    print(f"{k:12s} {v['ms']:9.4f} ms  {v['achieved_gbs']:8.1f} GB/s  {v['frac_of_measured_peak']:.4f}")
print(d['config'])
if 'vehicle_step' in d:
    v = d['vehicle_step']
    print('vehicle', v['value'], v['ms_per_tick'], v['roofline']['frac'], v['e2e']['value'])
