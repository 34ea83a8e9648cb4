"""Key limiter metrics of every kernel in an ncu report: `python scripts/ncu_kernel_summary.py X.ncu-rep [regex]`."""
import csv, subprocess, sys, io
args = ["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"]
if len(sys.argv) > 2:
    args += ["--kernel-name", "regex:" + sys.argv[2]]
out = subprocess.run(args, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
KEYS = [
    ("time_us", "gpu__time_duration.sum"),
    ("regs", "launch__registers_per_thread"),
    ("smem_KB/blk", "launch__shared_mem_per_block_allocated"),
    ("warps_active%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue_active%", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
    ("lsu_wavefronts%", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    ("smem_instr", "smsp__sass_inst_executed_op_shared.sum"),
    ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
    ("bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
    ("warp_instr", "smsp__inst_executed.sum"),
    ("dram_read_MB", "dram__bytes_read.sum"),
    ("dram_write_MB", "dram__bytes_write.sum"),
    ("dram_TB/s", "dram__bytes.sum.per_second"),
    ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("lat/inst", "smsp__average_warp_latency_per_inst_issued.ratio"),
    ("st_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
    ("st_short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
    ("st_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("st_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("st_mio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
    ("st_lg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"),
    ("st_branch", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"),
    ("st_notsel", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
    ("st_sleep", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"),
    ("st_membar", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"),
]
for r in rows[2:]:
    print("==", r[ix["Kernel Name"]][:90], "grid", r[ix.get("Grid Size", 0)] if "Grid Size" in ix else "")
    print("   " + "  ".join(f"{n}={r[ix[k]]}" for n, k in KEYS if k in ix))
