#!/usr/bin/env python
"""Per-region view of an ncu source page: executed warp-instructions and stall samples in chunks of
N SASS instructions, with the dominant opcodes of each chunk.  Usage:
python profiles/ncu_regions.py x.ncu-rep [chunk]"""
import csv, io, subprocess, sys
from collections import Counter
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
CH = int(sys.argv[2]) if len(sys.argv) > 2 else 200
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ins = [(r[isrc], int(r[isamp] or 0), int(r[iex] or 0), r) for r in data]
ts, te = sum(x[1] for x in ins), sum(x[2] for x in ins)
print(f"{len(ins)} SASS instructions, {te} executed warp-instructions, {ts} samples")
def opc(t):
    t = t.split()
    t = t[1] if t[0].startswith("@") else t[0]
    return t.split(".")[0]
for i in range(0, len(ins), CH):
    ch = ins[i:i + CH]
    s, e = sum(x[1] for x in ch), sum(x[2] for x in ch)
    if e == 0 and s == 0: continue
    c = Counter(opc(x[0]) for x in ch)
    st = Counter()
    for x in ch:
        for j, h in stall_cols:
            v = x[3][j]
            if v: st[h[6:]] += int(v)
    top = ", ".join(f"{k}={v}" for k, v in st.most_common(4))
    print(f"{i:5d} exec {100*e/te:5.1f}% samp {100*s/ts:5.1f}%  {dict(c.most_common(4))}  stalls: {top}")
