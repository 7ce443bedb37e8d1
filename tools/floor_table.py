#!/usr/bin/env python
"""Per (entry point, shape) excess of the measured ncu launch time over a simple floor max(flop/peak, bytes/HBM, 2.5us).
usage: python tools/floor_table.py gpurun_out/launches_r01.csv"""
import collections, csv, re, sys
PEAK_TF, HBM = 1435.6e12, 6447.5e9
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0.0, 0.0, 0])
cat = collections.defaultdict(lambda: [0.0, 0.0, 0])
def shp(s):
    m = re.match(r'(\d+)x(\d+)x(\d+)x(\d+)([bf])', s)
    if not m: return None
    n, h, w, c = map(int, m.groups()[:4])
    return n * h * w, c, (2 if m.group(5) == 'b' else 4)
for r in csv.DictReader(lines):
    if r['Metric Name'] != 'gpu__time_duration.sum': continue
    t = float(r['Metric Value']) / 1e3
    lab, _, kern = r['Kernel Name'].partition('/')
    p = lab.split('|')
    name, args = p[1], p[2].split(',')
    sh = [shp(a) for a in args if shp(a)]
    ints = [int(a) for a in args if re.fullmatch(r'-?\d+', a)]
    flop = byt = 0.0
    if name in ('phs_conv2d', 'phs_conv2d_stats', 'phs_conv2d_stats_acc', 'phs_conv2d_wgrad') and len(sh) >= 2:
        k = ints[0]
        (px, c0, e0), (_, c1, e1) = sh[0], sh[1]
        flop = 2.0 * px * k * k * c0 * c1
        byt = px * (c0 * e0 + c1 * e1) + k * k * c0 * c1 * 2
    elif sh:
        passes = {'phs_norm_act_fwd': 2, 'phs_norm_bwd_reduce': 2, 'phs_norm_bwd_apply': 3, 'phs_chan_stats': 1}.get(name)
        if passes:
            px, c, e = sh[0]; byt = passes * px * c * e
        else:
            byt = sum(px * c * e for px, c, e in sh)
    floor = max(flop / PEAK_TF, byt / HBM, 2.5e-6) * 1e6
    key = name + '|' + p[2]
    agg[key][0] += t; agg[key][1] += floor; agg[key][2] += 1
    cat[name][0] += t; cat[name][1] += floor; cat[name][2] += 1
tt = sum(v[0] for v in cat.values()); tf = sum(v[1] for v in cat.values())
print('total %.0f us, floor %.0f us' % (tt, tf))
for k, (t, f, n) in sorted(cat.items(), key=lambda kv: -(kv[1][0] - kv[1][1])):
    print('%-28s n=%4d  time %8.1f  floor %8.1f  excess %8.1f' % (k, n, t, f, t - f))
print()
for k, (t, f, n) in sorted(agg.items(), key=lambda kv: -(kv[1][0] - kv[1][1]))[:50]:
    print('%-86s n=%3d  time %7.1f  floor %7.1f  excess %7.1f' % (k[:86], n, t, f, t - f))
