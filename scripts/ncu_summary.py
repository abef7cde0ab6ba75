"""Summarise an `ncu --set full` report of one kernel as markdown (profiles/*.md).

    python scripts/ncu_summary.py gpurun_out/r2ncu.ncu-rep "title" > profiles/r02_tc_summary.md
"""
import csv, io, subprocess, sys

rep, title = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu summary"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit = rows[0], rows[1]
WANT = [
    "gpu__time_duration.sum", "sm__cycles_active.avg", "launch__registers_per_thread", "launch__block_size",
    "launch__grid_size", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]
print(f"# {title}\n")
print(f"Source: `{rep}` (`ncu --set full --clock-control none --import-source on`, one launch).\n")
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(f"## {d.get('Kernel Name', '?')[:120]}\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in WANT:
        if k in d:
            print(f"| `{k}` | {d[k]} | {unit[hdr.index(k)]} |")
    print()
