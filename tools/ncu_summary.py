#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline discussion needs.
usage: ncu_summary.py report.ncu-rep [kernel-substring] > profiles/xxx.txt"""
import csv
import subprocess
import sys

KEYS = """gpu__time_duration.sum
sm__cycles_elapsed.max
sm__cycles_active.avg
launch__grid_size
launch__block_size
launch__registers_per_thread
launch__shared_mem_per_block_dynamic
launch__occupancy_limit_shared_mem
launch__occupancy_limit_registers
sm__warps_active.avg.per_cycle_active
sm__warps_active.avg.pct_of_peak_sustained_active
smsp__inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.sum
sm__inst_executed_pipe_xu.sum
sm__inst_executed_pipe_alu.sum
sm__inst_executed_pipe_fma.sum
sm__inst_executed_pipe_uniform.sum
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
smsp__inst_executed_op_global_ld.sum
smsp__inst_executed_op_global_st.sum
smsp__inst_executed_op_shared_ld.sum
smsp__inst_executed_op_shared_st.sum
l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum
l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum
l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum
l1tex__t_requests_pipe_lsu_mem_global_op_st.sum
l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum
l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum
l1tex__data_pipe_lsu_wavefronts.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed
l1tex__t_set_accesses.avg.pct_of_peak_sustained_elapsed
l1tex__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__t_sector_hit_rate.pct
lts__t_sectors_srcunit_tex_op_read.sum
lts__t_sectors_srcunit_tex_op_write.sum
lts__t_sector_hit_rate.pct
lts__throughput.avg.pct_of_peak_sustained_elapsed
lts__t_bytes.sum.per_second
dram__bytes_read.sum
dram__bytes_write.sum
dram__throughput.avg.pct_of_peak_sustained_elapsed
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
sm__throughput.avg.pct_of_peak_sustained_elapsed
smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio
smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio
smsp__average_warp_latency_issue_stalled_barrier.ratio
smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio
smsp__average_warp_latency_issue_stalled_lg_throttle.ratio
smsp__average_warp_latency_issue_stalled_mio_throttle.ratio
smsp__average_warp_latency_issue_stalled_wait.ratio
smsp__average_warp_latency_issue_stalled_not_selected.ratio
smsp__average_warp_latency_issue_stalled_branch_resolving.ratio
smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio
smsp__average_warp_latency_issue_stalled_no_instruction.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__warps_issue_stalled_long_scoreboard_per_warp_active.pct
smsp__warps_issue_stalled_barrier_per_warp_active.pct
smsp__warps_issue_stalled_short_scoreboard_per_warp_active.pct
smsp__warps_issue_stalled_mio_throttle_per_warp_active.pct
smsp__warps_issue_stalled_lg_throttle_per_warp_active.pct
smsp__warps_issue_stalled_wait_per_warp_active.pct
smsp__warps_issue_stalled_math_pipe_throttle_per_warp_active.pct
smsp__warps_issue_stalled_not_selected_per_warp_active.pct
smsp__warps_issue_stalled_selected_per_warp_active.pct
smsp__warps_issue_stalled_sleeping_per_warp_active.pct
smsp__warps_issue_stalled_membar_per_warp_active.pct
smsp__warps_issue_stalled_tex_throttle_per_warp_active.pct
smsp__warps_issue_stalled_branch_resolving_per_warp_active.pct
smsp__warps_issue_stalled_dispatch_stall_per_warp_active.pct
smsp__warps_issue_stalled_drain_per_warp_active.pct
smsp__warps_issue_stalled_imc_miss_per_warp_active.pct
smsp__warps_issue_stalled_no_instruction_per_warp_active.pct
smsp__thread_inst_executed_per_inst_executed.ratio""".split()


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if sub and sub not in name:
            continue
        print("== %s  grid %s block %s  (id %s)" % (name, r[col.get("Grid Size", 0)], r[col.get("Block Size", 0)], r[0]))
        for k in KEYS:
            if k in col and r[col[k]] != "":
                print("  %-82s %14s %s" % (k, r[col[k]], units[col[k]]))


if __name__ == "__main__":
    main()
