"""Summarise one kernel of an `ncu --set full` report into profiles/ncu_summary.json (machine-readable, read by bench.py for
roofline.traffic) and append a table to a markdown file under profiles/.

usage: python scripts/ncu_summary.py <key> <report.ncu-rep> <markdown-out> "<title>" "<command that produced the report>"
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__icc_request_hit_rate.pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__pcsamp_warps_issue_stalled_wait",
    "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_not_selected",
    "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_warps_issue_stalled_no_instructions", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
    "smsp__pcsamp_warps_issue_stalled_mio_throttle", "gpc__cycles_elapsed.avg.per_second",
]
UNIT_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    key, rep, md, title, command = sys.argv[1:6]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    out = {"kernel": vals[col["Kernel Name"]], "command": command}
    lines = [f"\n## {title}\ncommand: `{command}`; kernel `{out['kernel']}`; report `{os.path.basename(rep)}` (scratch, not committed)\n",
             "| metric | unit | value (one captured launch) |", "|---|---|---|"]
    for m in WANT:
        if m in col:
            out[m] = vals[col[m]]
            out[m + ".unit"] = units[col[m]]
            lines.append(f"| {m} | {units[col[m]]} | {vals[col[m]]} |")
    rd = float(out.get("dram__bytes_read.sum", 0) or 0) * UNIT_BYTES.get(out.get("dram__bytes_read.sum.unit", "byte"), 1.0)
    wr = float(out.get("dram__bytes_write.sum", 0) or 0) * UNIT_BYTES.get(out.get("dram__bytes_write.sum.unit", "byte"), 1.0)
    out["dram_bytes_per_launch"] = rd + wr
    lines.append(f"| dram bytes per launch (read + write) | byte | {rd + wr:.0f} |")
    path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[key] = out
    json.dump(data, open(path, "w"), indent=1)
    with open(md, "a") as fh:
        fh.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
