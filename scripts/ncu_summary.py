"""Summarise an .ncu-rep (ncu --set full) into a small markdown file for profiles/.

usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name.md "title" [events_per_launch] [tiles_per_warp]
"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
events = float(sys.argv[4]) if len(sys.argv) > 4 else None
tiles = float(sys.argv[5]) if len(sys.argv) > 5 else None

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[-1]
get = lambda k: vals[hdr.index(k)] if k in hdr else "n/a"
keys = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed_op_shared_atom.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]
lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none --import-source on), one launch", "",
         "| metric | value | unit |", "|---|---|---|"]
for k in keys:
    if k in hdr:
        lines.append(f"| {k} | {get(k)} | {units[hdr.index(k)]} |")
inst = float(get("smsp__inst_executed.sum").replace(",", ""))
if events:
    lines += ["", f"events in this launch: {events:.6g}; **warp-instructions per event: {inst / events:.1f}**"]
    if tiles:
        lines.append(f"replicates per warp: {tiles:g}; warp-instructions per warp-iteration: {inst / events * tiles:.0f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 3:
    h = rows[1]
    ia, isrc, isamp = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
    data = [r for r in rows[2:] if len(r) > ia and r[ia].isdigit()]
    tot = sum(int(r[ia]) for r in data)
    c, cs = Counter(), Counter()
    for r in data:
        parts = r[isrc].split()
        op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
        c[op] += int(r[ia])
        cs[op] += int(r[isamp])
    lines += ["", "## executed instructions by opcode (SASS, top 16)", "", "| opcode | share of executed | share of stall samples |",
              "|---|---|---|"]
    stot = sum(cs.values()) or 1
    for op, v in c.most_common(16):
        lines.append(f"| {op} | {100 * v / tot:.1f}% | {100 * cs[op] / stot:.1f}% |")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
