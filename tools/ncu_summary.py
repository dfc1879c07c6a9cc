#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md / profiles/ quote."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum',
 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','dram__throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__block_size','launch__grid_size',
 'sm__warps_active.avg.pct_of_peak_sustained_active','launch__occupancy_limit_registers',
 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
 'l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__inst_executed.sum','smsp__inst_executed_pipe_fp64.sum',
 'sm__inst_executed_pipe_fp64.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
 'smsp__issue_active.avg.pct_of_peak_sustained_active','lts__t_bytes.sum','l1tex__t_bytes.sum','sm__cycles_elapsed.avg.per_second',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
 'smsp__average_warp_latency_issue_stalled_long_scoreboard.pct','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
 'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct','smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct',
 'smsp__warp_issue_stalled_wait_per_warp_active.pct','smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
 'smsp__warp_issue_stalled_not_selected_per_warp_active.pct','smsp__warp_issue_stalled_no_instruction_per_warp_active.pct',
 'smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct','smsp__warp_issue_stalled_drain_per_warp_active.pct',
 'smsp__warp_issue_stalled_imc_miss_per_warp_active.pct','smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct']
def main(path, grep=None):
    out = subprocess.run(['ncu','-i',path,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for val in rows[2:]:
        name = val[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''
        print('# kernel:', name)
        for i,h in enumerate(hdr):
            if h in WANT or (grep and grep in h):
                print(f'{h:80s} {val[i]:>20s} {units[i]}')
if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv)>2 else None)
