#!/usr/bin/env python
"""Summarise an .ncu-rep (first profiled kernel): key raw metrics + executed-instruction histogram
by opcode + top stall reasons.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv, io, subprocess, sys
from collections import Counter

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg", "smsp__thread_inst_executed_per_inst_executed.ratio"]
for h, u, v in zip(hdr, units, vals):
    if h in keys:
        print(f"{h} [{u}] = {v}")
stalls = sorted(((float(v), h) for h, v in zip(hdr, vals) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")), reverse=True)
print("stalls per issue:", ", ".join(f"{h.split('stalled_')[1].split('_per')[0]}={x:.2f}" for x, h in stalls[:7]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]; ci = {h: i for i, h in enumerate(h2)}
data = rows[2:]
c, s = Counter(), Counter()
for r in data:
    t = r[ci["Source"]].strip().split()
    if not t: continue
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    c[op] += int(r[ci["Instructions Executed"]] or 0); s[op] += int(r[ci["# Samples"]] or 0)
tot, ts = sum(c.values()), max(1, sum(s.values()))
print(f"executed warp-instructions: {tot}")
for op, n in c.most_common(16):
    print(f"  {op:8s} {100*n/tot:5.1f}% of instructions, {100*s[op]/ts:5.1f}% of stall samples")

# optional: python tools/ncu_summary.py REP --traffic-json KEY  -> merges the capture's DRAM bytes per launch into
# profiles/ncu_traffic.json under KEY (bench.py's roofline.traffic reads it for the matching workload)
if len(sys.argv) > 3 and sys.argv[2] == "--traffic-json":
    import json, os
    key = sys.argv[3]
    get = lambda name: float(vals[hdr.index(name)].replace(",", "")) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[units[hdr.index(name)]]
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[key] = {"dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
                 "kernel": vals[hdr.index("Kernel Name")], "duration_ms_under_ncu": vals[hdr.index("gpu__time_duration.sum")],
                 "source": "ncu --set full --clock-control none, one launch: " + os.path.basename(rep)}
    json.dump(data, open(path, "w"), indent=1)
    print("wrote", path, key, data[key])
