#!/usr/bin/env python
"""Text summary of an .ncu-rep (run here, no GPU needed): the metrics the roofline arithmetic uses, per profiled launch.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [label per launch, comma separated] > profiles/x.txt"""
import csv
import subprocess
import sys

METRICS = [
    ('gpu__time_duration.sum', 'duration'),
    ('sm__cycles_elapsed.avg.per_second', 'SM clock'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__registers_per_thread', 'regs/thread'),
    ('launch__shared_mem_per_block_dynamic', 'dyn smem/block'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe active'),
    ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit rate'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active'),
    ('smsp__inst_executed.sum', 'warp instructions'),
]


def main():
    rep = sys.argv[1]
    labels = sys.argv[2].split(',') if len(sys.argv) > 2 else []
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn = hdr.index('Kernel Name')
    print('# %s  (ncu --set full --clock-control none --import-source on; cold-cache, serialised replays)' % rep)
    for i, d in enumerate(data):
        lab = labels[i] if i < len(labels) else ''
        print('\n[%d] %s  %s' % (i, d[kn][:100], lab))
        for m, nice in METRICS:
            if m in hdr:
                j = hdr.index(m)
                print('    %-22s %14s %s' % (nice, d[j], units[j]))


if __name__ == '__main__':
    main()
