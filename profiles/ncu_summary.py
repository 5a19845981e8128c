#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the handful of metrics we track (one row per captured launch)."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active']


def main(rep, out):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['kernel'] + [f'{k} [{units[ix[k]]}]' for k in KEYS if k in ix])
        for r in data:
            w.writerow([r[ix['Kernel Name']][:80]] + [r[ix[k]] for k in KEYS if k in ix])


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
