#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed): python tools/ncu_summary.py report.ncu-rep [more keys...]"""
import csv
import subprocess
import sys

KEYS = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'launch__waves_per_multiprocessor', 'launch__shared_mem_per_block_dynamic', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_barrier_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_wait_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_membar_per_warp_active.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def main():
    rep = sys.argv[1]
    keys = KEYS + sys.argv[2:]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('-' * 100)
        for k in keys:
            if k in hdr:
                i = hdr.index(k)
                print(f'{k:80s} {r[i][:60]:>24s} {units[i]}')


if __name__ == '__main__':
    main()
