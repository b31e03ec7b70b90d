#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full, brought back in gpurun_out/) into the small text summary kept
under profiles/:  python profiles/summarize.py gpurun_out/prof.ncu-rep profiles/name.txt [series_per_launch]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
nser = float(sys.argv[3]) if len(sys.argv) > 3 else 1e6
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
lines = ["# %s" % rep]
for i, h in enumerate(hdr):
    if h in keep or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
        try:
            if h != "Kernel Name" and float(vals[i].replace(",", "")) == 0:
                continue
        except ValueError:
            pass
        lines.append("%-80s %-12s %s" % (h, units[i], vals[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2 and "Source" in rows[1]:
    h = rows[1]
    si, ei, sm = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    ops, samp = collections.Counter(), collections.Counter()
    for r in rows[2:]:
        try:
            n, s = float(r[ei].replace(",", "")), float(r[sm].replace(",", ""))
        except (ValueError, IndexError):
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[si])
        if m:
            op = m.group(2).split(".")[0]
            ops[op] += n
            samp[op] += s
    tot, ts = sum(ops.values()), max(1.0, sum(samp.values()))
    lines.append("# SASS opcode mix (warp instructions per series, %% of instructions, %% of stall samples); total %.0f per series" % (tot / nser))
    for k, v in ops.most_common(24):
        lines.append("%-12s %9.1f %6.1f%% %6.1f%%" % (k, v / nser, 100 * v / tot, 100 * samp[k] / ts))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:12]))
