#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU): headline metrics, pipe utilisation, stall breakdown per opcode.
usage: tools/ncu_summary.py report.ncu-rep"""
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
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
print("kernel:", m.get("Kernel Name", ("?",))[0])
for k in keys:
    if k in m:
        print(f"  {k:75s} {m[k][0]:>16s} {m[k][1]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter(); byop = collections.defaultdict(collections.Counter); cnt = collections.Counter(); execd = collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr): continue
    mm = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip())
    op = mm.group(2) if mm else "?"
    n = int(r[ix["# Samples"]]); cnt[op] += n; execd[op] += int(r[ix["Instructions Executed"]])
    for s in stalls:
        v = int(r[ix[s]]); tot[s] += v; byop[op][s] += v
T = sum(cnt.values()) or 1
print("warp-instructions executed:", sum(execd.values()))
print("stall samples:", ", ".join(f"{s[6:]} {100 * v / T:.1f}%" for s, v in tot.most_common(8)))
for op, n in cnt.most_common(14):
    print(f"  {op:24s} samples {100 * n / T:5.1f}%  executed {execd[op]:11d}  " + ", ".join(f"{s[6:]}={v}" for s, v in byop[op].most_common(3)))
