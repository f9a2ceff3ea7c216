#!/usr/bin/env python
"""Summarise an .ncu-rep (first profiled launch) into the markdown kept under profiles/.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md"""
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
get = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__cycles_active.avg", "sm__cycles_elapsed.max",
]
print(f"# ncu summary: {rep.split('/')[-1]}\n")
print(f"kernel: `{get.get('Kernel Name', ('?', ''))[0]}`\n")
print("| metric | value | unit |\n|---|---|---|")
for w in WANT:
    if w in get:
        print(f"| {w} | {get[w][0]} | {get[w][1]} |")
try:
    rd, wr = float(get["dram__bytes_read.sum"][0]), float(get["dram__bytes_write.sum"][0])
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
    tot = rd * scale.get(get["dram__bytes_read.sum"][1], 1) + wr * scale.get(get["dram__bytes_write.sum"][1], 1)
    print(f"\nDRAM traffic (read+write) per launch: **{tot / 1e6:.1f} MB**")
except (KeyError, ValueError):
    pass

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
if len(srows) > 3:
    sh, data = srows[1], srows[2:]
    ia, ie = sh.index("Source"), sh.index("Instructions Executed")
    st = [i for i, h in enumerate(sh) if h.startswith("stall_") and "Not Issued" not in h]
    ops, stalls, tot = collections.Counter(), collections.Counter(), 0
    for r in data:
        try:
            n = int(r[ie])
        except ValueError:
            continue
        tot += n
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia])
        ops[m.group(2).split(".")[0] if m else "?"] += n
        for i in st:
            try:
                stalls[sh[i][6:]] += int(r[i])
            except ValueError:
                pass
    print("\nexecuted warp instructions by opcode (top 16): " +
          ", ".join(f"{k} {v / tot * 100:.1f}%" for k, v in ops.most_common(16)))
    f64 = sum(v for k, v in ops.items() if k in ("DMUL", "DADD", "DFMA", "DSETP"))
    print(f"\nf64 share of executed instructions: {f64 / tot * 100:.1f}%")
    s = sum(stalls.values())
    if s:
        print("\nwarp stall samples: " + ", ".join(f"{k} {v / s * 100:.1f}%" for k, v in stalls.most_common(8)))
