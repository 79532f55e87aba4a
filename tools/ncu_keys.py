"""Prints the metrics that decide what bounds a kernel from an `ncu --page raw --csv` export (stdin or file)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.max.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
        "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
for r in rows[2:]:
    print("Kernel Name".ljust(88), r[hdr.index("Kernel Name")][:120])
    for i, h in enumerate(hdr):
        if h in want or ("issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h):
            print(f"{h:88s} {units[i]:16s} {r[i]}")
