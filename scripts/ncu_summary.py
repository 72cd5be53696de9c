#!/usr/bin/env python3
"""Summarise an ncu report (--set full, --import-source on) of the solve kernel into markdown.
usage: scripts/ncu_summary.py gpurun_out/prof_r1.ncu-rep > profiles/solve_kernel_r1_summary.md"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
M = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
want = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__warps_eligible.avg.per_cycle_active",
]
print(f"# ncu summary of `{M.get('Kernel Name', ('solve_kernel', ''))[0][:60]}` ({rep})\n")
print("Captured with `ncu --set full --clock-control none --import-source on -k regex:solve_kernel` around "
      "`bench.py --steps 2 --warmup 1` (config C3: 65536 problems, control_steps 10).  Times under the profiler are "
      "cold-cache and serialised; bench values come from un-profiled runs.\n")
print("| metric | value | unit |\n|---|---|---|")
for k in want:
    if k in M:
        print(f"| {k} | {M[k][0]} | {M[k][1]} |")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur = None; h = None; agg = {}; ops = collections.Counter(); tot_ops = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        h = r; ie = h.index("Instructions Executed"); isamp = h.index("# Samples"); continue
    if h is None or len(r) <= ie:
        continue
    if r[0].isdigit() and r[2] == "-":
        try:
            agg[(cur, int(r[0]))] = (int(r[ie]), int(r[isamp]), r[1])
        except ValueError:
            pass
    elif r[2].startswith("0x") or (len(r) > 3 and r[0] == "" ):
        pass
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(sass)))
sh = srows[1]; ia = sh.index("Source"); ie2 = sh.index("Instructions Executed")
for r in srows[2:]:
    if len(r) <= ie2:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia])
    op = m.group(2).split(".")[0] if m else "?"
    n = int(r[ie2]); ops[op] += n; tot_ops += n
print(f"\nStatic SASS instructions: {len(srows) - 2}; executed warp instructions: {tot_ops}\n")
print("## Instruction mix (executed warp instructions)\n\n| opcode | share |\n|---|---|")
for op, n in ops.most_common(16):
    print(f"| {op} | {100 * n / tot_ops:.1f} % |")
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print("\n## Hottest source lines (share of executed instructions / of stall samples)\n\n| inst | samples | line | source |\n|---|---|---|---|")
for (f, l), (n, s, text) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    t = text.strip().replace("|", "\\|")[:90]
    print(f"| {100 * n / tot:.2f} % | {100 * s / max(tots, 1):.2f} % | {f}:{l} | `{t}` |")
