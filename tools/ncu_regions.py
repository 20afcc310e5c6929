#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + stall samples aggregated over SASS regions (no GPU needed)."""
import csv, io, re, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum.per_cycle_elapsed", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__block_size"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:75s} {units[i]:12s} {[r[i] for r in data]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[start]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for r in rows[start + 1:]:
    if r and r[0] == "Address":
        break
    try:
        data.append((int(r[isamp] or 0), int(r[iex] or 0), r[ia]))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print("samples", tot, "instructions", len(data))
W = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for b in range(0, len(data), W):
    chunk = data[b:b + W]
    s = sum(c[0] for c in chunk)
    if s < tot * 0.015:
        continue
    ops = {}
    for c in chunk:
        m = re.search(r"([A-Z][A-Z0-9_]+)", re.sub(r"@!?U?P\d+", "", c[2]))
        if m:
            ops[m.group(1)] = ops.get(m.group(1), 0) + c[0]
    top = sorted(ops.items(), key=lambda x: -x[1])[:6]
    print(f"{b:5d} {100 * s / tot:5.1f}%  exec {max(c[1] for c in chunk):8d}  {top}")
