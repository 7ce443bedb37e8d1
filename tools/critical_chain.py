#!/usr/bin/env python
"""Offline critical-chain estimate from a tools/timeline.py chrome trace.  Walks back from the last device activity of
the step: the predecessor of an activity is the one (any stream) that ended last before it started; time on the chain is
attributed to kernel families, waits between a predecessor's end and the start are reported as gaps.  A heuristic (the
trace has no dependency edges), good enough to see WHICH kernels the step's length is made of.
usage: python tools/critical_chain.py gpurun_out/r2/timeline_r02_chrome.json.gz"""
import bisect, collections, gzip, json, re, sys

d = json.load(gzip.open(sys.argv[1]) if sys.argv[1].endswith('.gz') else open(sys.argv[1]))
ev = [e for e in d['traceEvents'] if e.get('cat') in ('kernel', 'gpu_memset', 'gpu_memcpy')]
ev.sort(key=lambda e: e['ts'])
h2d = [i for i, e in enumerate(ev) if e['name'].startswith('Memcpy HtoD')]
start = h2d[-1] if h2d else 0
step = ev[start:]
def fam(n):
    n = re.sub(r'^void ', '', n); n = re.sub(r'\(anonymous namespace\)::', '', n); n = re.sub(r'<unnamed>::', '', n); n = re.sub(r'[<(].*', '', n)
    return n[:36]
ends = sorted((e['ts'] + e['dur'], i) for i, e in enumerate(step))
end_t = [t for t, _ in ends]
cur = max(range(len(step)), key=lambda i: step[i]['ts'] + step[i]['dur'])
on = collections.Counter(); cnt = collections.Counter(); gap = 0.0; ngap = 0; chain = []
while True:
    e = step[cur]
    chain.append(cur)
    k = bisect.bisect_right(end_t, e['ts'] + 0.5) - 1      # last activity that ended before this one started
    if k < 0:
        on[fam(e['name'])] += e['dur']; cnt[fam(e['name'])] += 1
        break
    pred = ends[k][1]
    if pred == cur:
        k -= 1
        if k < 0: break
        pred = ends[k][1]
    pe = step[pred]['ts'] + step[pred]['dur']
    on[fam(e['name'])] += e['ts'] + e['dur'] - max(pe, e['ts']) if pe > e['ts'] else e['dur']
    cnt[fam(e['name'])] += 1
    if pe < e['ts']:
        gap += e['ts'] - pe; ngap += 1
    cur = pred
tot = step[chain[0]]['ts'] + step[chain[0]]['dur'] - step[chain[-1]]['ts']
print('chain of %d activities spans %.0f us; gaps between a predecessor and its successor: %d, %.0f us' % (len(chain), tot, ngap, gap))
print('%-38s %8s %8s' % ('kernel family on the chain', 'count', 'us'))
for f, v in on.most_common():
    print('%-38s %8d %8.0f' % (f, cnt[f], v))
