#!/usr/bin/env python
"""Print the metrics we track from an .ncu-rep (needs `ncu` on PATH; no GPU required).
usage: tools/ncu_summary.py <report.ncu-rep> [--md]"""
import csv, io, subprocess, sys

WANT = ['gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__occupancy_limit_registers',
 'dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'lts__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','lts__t_sectors_srcunit_tex_op_read.sum',
 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__t_sector_hit_rate.pct','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active',
 'smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_tensor.sum','sm__cycles_elapsed.max',
 'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio','smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio']

def main():
    rep = sys.argv[1]
    md = '--md' in sys.argv
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')].split('(')[0].split('::')[-1]
        print(('### ' if md else '--- ') + name)
        if md:
            print('\n| metric | value | unit |\n|---|---:|---|')
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f'| {w} | {r[i]} | {units[i]} |' if md else f'{w:85s} {r[i]:>18s} {units[i]}')

main()
