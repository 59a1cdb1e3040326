"""Condense an `ncu --set full` report into one JSON record per kernel launch.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_full_summary.json   (needs ncu on PATH, no GPU)"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed": "fmaheavy_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_inst_pct",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active": "fmaheavy_inst_pct",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_pct",
    "sm__inst_issued.avg.per_cycle_active": "ipc",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_inst_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_per_inst",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
}


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {name: i for i, name in enumerate(hdr)}
    out = []
    for r in data:
        rec = {"kernel": r[col["Kernel Name"]][:70]}
        for metric, key in WANT.items():
            if metric in col:
                rec[key] = f"{r[col[metric]]} {units[col[metric]]}".strip()
        out.append(rec)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
