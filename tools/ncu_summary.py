#!/usr/bin/env python
"""Summarises an .ncu-rep (read here, no GPU): the metrics B200_PROFILING.md asks for, per profiled launch."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps',
        'launch__waves_per_multiprocessor', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'derived__smsp__sass_thread_inst_executed_op_memory_pred_on.sum', 'local_load', 'smsp__sass_inst_executed_op_local_ld.sum',
        'smsp__sass_inst_executed_op_local_st.sum']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('### kernel:', r[hdr.index('Kernel Name')][:90], ' grid', r[hdr.index('Grid Size')], 'block', r[hdr.index('Block Size')])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f'{k:86s} {r[i]:>16s} {units[i]}')


if __name__ == '__main__':
    main(sys.argv[1])
