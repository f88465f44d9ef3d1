#!/usr/bin/env python
"""Text summary of one kernel from an .ncu-rep (read with `ncu -i`): the metrics the roofline uses,
the warp-stall breakdown and the SASS opcode mix of the hottest loop.  usage: ncu_summary.py REP"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units, row = r[0], r[1], r[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sass__inst_executed_local_loads",
        "sass__inst_executed_local_stores", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
print(f"# {rep}")
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:72s} {units[i]:14s} {row[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
if len(rows) > 2:
    h = rows[1]
    ix = {k: i for i, k in enumerate(h)}
    data = [rr for rr in rows[2:] if len(rr) >= len(h)]
    st = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    tot = collections.Counter()
    ns = 0
    for rr in data:
        ns += int(rr[ix["# Samples"]] or 0)
        for k in st:
            tot[k] += int(rr[ix[k]] or 0)
    print(f"\nwarp-stall samples: {ns}")
    for k, v in tot.most_common(10):
        print(f"  {k:28s} {100 * v / max(ns, 1):5.1f}%")
    cnt = collections.Counter(int(rr[ix["Instructions Executed"]]) for rr in data)
    big = [c for c, n in cnt.items() if n > 40 and c > 0]
    if big:
        loopc = max(big)
        mix = collections.Counter()
        nins = 0
        for rr in data:
            if int(rr[ix["Instructions Executed"]]) >= 0.45 * loopc:
                op = rr[ix["Source"]].strip()
                if op.startswith("@"):
                    op = op.split(None, 1)[1]
                mix[op.split()[0].split(".")[0]] += 1
                nins += 1
        print(f"\nhottest loop: {nins} SASS instructions per iteration, executed {loopc} times per instruction (warp-level)")
        print("  " + ", ".join(f"{k} {v}" for k, v in mix.most_common(24)))
