#!/usr/bin/env python
"""Where a tile kernel's warp-samples fall: inside the barrier-carrying time loops or in the per-tile prologue / epilogue.
Reads the SASS source page of an .ncu-rep (first kernel) -- usage: python tools/ncu_regions.py REP [top_n]"""
import csv, io, re, subprocess, sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
h = rows[hi]; ci = {k: i for i, k in enumerate(h)}
data = [r for r in rows[hi + 1:] if len(r) == len(h)]
ins = []
for r in data:
    a = r[ci["Address"]] if "Address" in ci else ""
    ins.append(dict(addr=int(a, 16) if a else len(ins) * 16, sass=r[ci["Source"]].strip(), samples=int(r[ci["# Samples"]] or 0),
                    execd=int(r[ci["Instructions Executed"]] or 0), row=r))
base = ins[0]["addr"]
for x in ins: x["addr"] -= base
idx = {x["addr"]: i for i, x in enumerate(ins)}
in_loop = [False] * len(ins)
loops = []
for i, x in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?`?\(?0x([0-9a-f]+)", x["sass"])
    if m:
        tgt = int(m.group(1), 16) - (base if int(m.group(1), 16) >= base else 0)
        if tgt < x["addr"] and tgt in idx:
            body = ins[idx[tgt]: i + 1]
            nbar = sum("BAR" in b["sass"] for b in body)
            if nbar >= 2 and len(body) < 1500:
                loops.append((tgt, x["addr"], len(body)))
                for k in range(idx[tgt], i + 1): in_loop[k] = True
tot = sum(x["samples"] for x in ins)
lp = sum(x["samples"] for x, f in zip(ins, in_loop) if f)
print(f"{len(ins)} SASS instructions, {len(loops)} time loops; samples: total {tot}, in time loops {lp} ({100*lp/max(tot,1):.1f}%), outside {tot-lp} ({100*(tot-lp)/max(tot,1):.1f}%)")
ex_l = sum(x["execd"] for x, f in zip(ins, in_loop) if f); ex = sum(x["execd"] for x in ins)
print(f"warp-instructions executed: total {ex}, in loops {ex_l} ({100*ex_l/max(ex,1):.1f}%)")
out = sorted((x for x, f in zip(ins, in_loop) if not f), key=lambda x: -x["samples"])[:top_n]
stall_cols = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
for name, sel in (("in loops", True), ("outside", False)):
    agg = {k: sum(int(x["row"][ci[k]] or 0) for x, f in zip(ins, in_loop) if f == sel) for k in stall_cols}
    tt = max(1, sum(agg.values()))
    print(f"stall reasons {name}: " + ", ".join(f"{k[6:]}={100*v/tt:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for x in out:
    print(f"  {x['addr']:#7x} {x['samples']:7d} ({100*x['samples']/max(tot,1):4.1f}%) x{x['execd']:<9d} {x['sass'][:90]}")
