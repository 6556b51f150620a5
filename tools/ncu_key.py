"""Key metrics from an `ncu --page raw --csv` dump."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "sm__cycles_elapsed.avg.per_second", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio"]
for i, h in enumerate(hdr):
    if h in KEYS or ("issue_stalled" in h and h.endswith("_per_warp_active.pct")) or ("warps_issue_stalled" in h and "pct" in h):
        print(f"{h:90s} {units[i]:12s} {vals[i]}")
