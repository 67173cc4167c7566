#!/usr/bin/env python
"""List the barrier-carrying loops of one kernel in a built library with their instruction mix
(cuobjdump -sass; no GPU needed).  usage: sass_loops.py LIB.so 'k_tileILi1ELb1ELi2ENS_5Exact'"""
import re
import subprocess
import sys
from collections import Counter

lib, key = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, L = None, []
for line in txt.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    if fn and key in fn:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*;", line)
        if m:
            L.append((int(m.group(1), 16), m.group(2)))
addr = {a: i for i, (a, _) in enumerate(L)}
print(len(L), "instructions")
for i, (a, ins) in enumerate(L):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?`?\(?0x([0-9a-f]+)", ins)
    if m and int(m.group(1), 16) < a and int(m.group(1), 16) in addr:
        body = [x[1] for x in L[addr[int(m.group(1), 16)]: i + 1]]
        ops = Counter(re.sub(r"^@!?U?P\d\s+", "", b).split()[0].split(".")[0] for b in body)
        if ops.get("BAR", 0) >= 2:
            dp = ops["DADD"] + ops["DMUL"] + ops["DFMA"]
            print(f"loop {int(m.group(1),16):#x}..{a:#x}: {len(body)} instr, {dp} fp64, {len(body)-dp} other:", dict(ops.most_common()))
            if len(sys.argv) > 3 and sys.argv[3] == hex(int(m.group(1), 16)):
                print("\n".join(body))
