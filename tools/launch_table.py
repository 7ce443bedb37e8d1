#!/usr/bin/env python
"""Aggregate an NVTX-labelled ncu launch list (tools/ncu_step.py) per kernel and per (entry point, shape).
usage: python tools/launch_table.py gpurun_out/launches_r01.csv [top]"""
import collections, csv, re, sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
per_kernel = collections.defaultdict(lambda: [0.0, 0])
per_shape = collections.defaultdict(lambda: [0.0, 0, ''])
total = 0.0
for r in rows:
    if r['Metric Name'] != 'gpu__time_duration.sum':
        continue
    t = float(r['Metric Value']) / 1e3
    kn = r['Kernel Name']
    lab, _, kern = kn.partition('/')
    kern = re.sub(r'\(.*', '', kern)
    parts = lab.split('|')
    key = '|'.join(parts[1:]) if len(parts) > 1 else lab
    per_kernel[kern][0] += t; per_kernel[kern][1] += 1
    per_shape[key + ' :: ' + kern][0] += t; per_shape[key + ' :: ' + kern][1] += 1
    total += t
print('# total %.1f us over %d launches' % (total, sum(v[1] for v in per_kernel.values())))
print('%-60s %8s %10s %6s' % ('kernel', 'launches', 'time_us', 'share'))
for k, (t, n) in sorted(per_kernel.items(), key=lambda kv: -kv[1][0]):
    print('%-60s %8d %10.1f %5.1f%%' % (k[:60], n, t, 100 * t / total))
print()
print('%-110s %4s %9s %6s %8s' % ('entry point | shapes :: kernel', 'n', 'time_us', 'share', 'us/launch'))
for k, (t, n, _) in sorted(per_shape.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%-110s %4d %9.1f %5.1f%% %8.1f' % (k[:110], n, t, 100 * t / total, t / n))
