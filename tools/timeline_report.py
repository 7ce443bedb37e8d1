#!/usr/bin/env python
"""Offline report of a tools/timeline.py chrome trace: step span, GPU busy time, concurrency histogram, and the time
attributable to each kernel family when it runs ALONE on the GPU (nothing else in flight) vs overlapped.
usage: python tools/timeline_report.py gpurun_out/timeline_chrome.json.gz"""
import collections, gzip, json, re, sys

d = json.load(gzip.open(sys.argv[1]) if sys.argv[1].endswith('.gz') else open(sys.argv[1]))
ev = [e for e in d['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memset', 'gpu_memcpy')]
ev.sort(key=lambda e: e['ts'])
# the last step: everything after the last HtoD copy burst
h2d = [i for i, e in enumerate(ev) if e['name'].startswith('Memcpy HtoD')]
start = h2d[-1] if h2d else 0
while start > 0 and ev[start - 1]['name'].startswith('Memcpy HtoD'):
    start -= 1
step = ev[start:]
t0 = step[0]['ts']
t1 = max(e['ts'] + e['dur'] for e in step)
print('step span %.1f us, %d device activities' % (t1 - t0, len(step)))
# sweep line
pts = []
for i, e in enumerate(step):
    pts.append((e['ts'], 1, i)); pts.append((e['ts'] + e['dur'], -1, i))
pts.sort()
active = set(); last = t0
conc = collections.Counter(); alone = collections.Counter(); shared = collections.Counter()
def fam(n):
    n = re.sub(r'^void ', '', n); n = re.sub(r'\(anonymous namespace\)::', '', n); n = re.sub(r'<.*', '', n)
    return n[:40]
for t, k, i in pts:
    dt = t - last
    if dt > 0:
        conc[len(active)] += dt
        if len(active) == 1:
            alone[fam(step[next(iter(active))]['name'])] += dt
        else:
            for j in active:
                shared[fam(step[j]['name'])] += dt / len(active)
    last = t
    if k == 1: active.add(i)
    else: active.discard(i)
tot = t1 - t0
print('concurrency (kernels in flight): ' + ', '.join('%d: %.0f us (%.0f%%)' % (k, v, 100 * v / tot) for k, v in sorted(conc.items())))
print('\n%-42s %10s %10s' % ('kernel family', 'alone us', 'shared us'))
fams = set(alone) | set(shared)
for f in sorted(fams, key=lambda f: -(alone[f] + shared[f])):
    print('%-42s %10.0f %10.0f' % (f, alone[f], shared[f]))
# idle gaps
gaps = []
cur_end = step[0]['ts'] + step[0]['dur']
for e in step[1:]:
    if e['ts'] > cur_end:
        gaps.append((e['ts'] - cur_end, cur_end - t0, e['name'][:50]))
    cur_end = max(cur_end, e['ts'] + e['dur'])
print('\nidle: %d gaps, %.0f us total; largest:' % (len(gaps), sum(g[0] for g in gaps)))
for g in sorted(gaps, reverse=True)[:8]:
    print('  %.1f us at t=%.0f before %s' % g)
# per-stream busy
bs = collections.defaultdict(float)
for e in step: bs[e['args'].get('stream')] += e['dur']
print('\nper-stream busy us:', {k: round(v) for k, v in bs.items()})
# coarse timeline: 0.5 ms buckets, dominant family + mean concurrency
print('\ntimeline (0.5 ms buckets): t_ms  mean-concurrency  top families by busy time')
B = 500.0
nb = int(tot / B) + 1
buck = [collections.Counter() for _ in range(nb)]
for e in step:
    a, b = e['ts'] - t0, e['ts'] - t0 + e['dur']
    k = int(a / B)
    while k * B < b and k < nb:
        lo, hi = max(a, k * B), min(b, (k + 1) * B)
        if hi > lo: buck[k][fam(e['name'])] += hi - lo
        k += 1
for k, c in enumerate(buck):
    s = sum(c.values())
    print('%5.1f  %4.2f  %s' % (k * B / 1e3, s / B, ', '.join('%s %.0f' % (f[:22], v) for f, v in c.most_common(4))))
