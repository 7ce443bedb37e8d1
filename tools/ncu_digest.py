#!/usr/bin/env python
"""Digest of an .ncu-rep (needs only the ncu CLI, no GPU): per profiled launch the roofline metrics, the pipe / memory
utilisation breakdown and the warp-stall reasons; with --source the hottest source lines (needs -lineinfo +
--import-source on).  usage: python tools/ncu_digest.py rep.ncu-rep [label,label,...] [--source N]"""
import csv, subprocess, sys

KEYS = [
    'gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
    'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_tensor.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'smsp__cycles_active.avg', 'sm__cycles_active.avg', 'l1tex__m_xbar2l1tex_read_bytes.sum',
    'l1tex__m_l1tex2xbar_write_bytes.sum', 'sm__memory_throughput.avg.pct_of_peak_sustained_elapsed',
]


def table(rep, page):
    raw = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(raw.splitlines()))


def main():
    rep = sys.argv[1]
    labels = sys.argv[2].split(',') if len(sys.argv) > 2 and not sys.argv[2].startswith('--') else []
    nsrc = int(sys.argv[sys.argv.index('--source') + 1]) if '--source' in sys.argv else 0
    rows = table(rep, 'raw')
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index('Kernel Name')
    print('# %s' % rep)
    for i, d in enumerate(data):
        print('\n[%d] %s  %s' % (i, d[kn][:90], labels[i] if i < len(labels) else ''))
        for k in KEYS:
            if k in hdr:
                j = hdr.index(k)
                print('    %-72s %16s %s' % (k, d[j], units[j]))
        stall = []
        for j, h in enumerate(hdr):
            if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
                try:
                    stall.append((float(d[j].replace(',', '')), h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')))
                except ValueError:
                    pass
        stall.sort(reverse=True)
        print('    warp stalls (warps per issue-active cycle): ' + ', '.join('%s %.2f' % (n, v) for v, n in stall[:8]))
    if nsrc:
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                             capture_output=True, text=True).stdout
        fpath, func, hdr = '', '', None
        per_func = {}
        for r in csv.reader(raw.splitlines()):
            if not r:
                continue
            if r[0] == 'File Path':
                fpath = r[1].split('/')[-1]
            elif r[0] == 'Function Name':
                func = r[1][:80]
            elif r[0] == 'Line No':
                hdr = r
            elif hdr and r[0].isdigit() and len(r) == len(hdr):
                ci = hdr.index('# Samples')
                try:
                    n = float(r[ci].replace(',', ''))
                except ValueError:
                    n = 0.0
                if n > 0:
                    key = (fpath, int(r[0]))
                    d = per_func.setdefault(func, {'n': {}, 'src': {}})
                    d['n'][key] = d['n'].get(key, 0.0) + n
                    d['src'][key] = r[1].strip()[:120]
        print('\n# hottest source lines (warp-stall samples, summed over the profiled launches of a kernel)')
        for func, d in per_func.items():
            tot = sum(d['n'].values()) or 1.0
            print('--- %s: %d samples' % (func, tot))
            for key, n in sorted(d['n'].items(), key=lambda kv: -kv[1])[:nsrc]:
                print('   %5.1f%%  %s:%d  %s' % (100 * n / tot, key[0], key[1], d['src'][key]))


if __name__ == '__main__':
    main()
