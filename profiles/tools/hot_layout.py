#!/usr/bin/env python
"""Hot-code footprint of a kernel from an `ncu --page source --csv` SASS export: how many distinct instructions most warps
execute and how they are spread over 2 KB (128-instruction) chunks -- the L1.5 instruction cache holds 32 KB.
usage: hot_layout.py <ncu_source.csv> [threshold_fraction]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ii = hdr.index("Instructions Executed")
recs = [(int(r[0], 16), int(r[ii] or 0)) for r in rows[hi + 1:] if len(r) > ii and r[0].startswith("0x")]
base = recs[0][0]
mx = recs[0][1]          # the first instruction is executed exactly once per warp
chunks = {}
for a, n in recs:
    c = (a - base) // 2048
    d = chunks.setdefault(c, [0, 0])
    d[0] += 1
    d[1] += n >= frac * mx
line = ""
for c in range(max(chunks) + 1):
    line += f"{chunks.get(c, [0, 0])[1]:4d}"
    if (c + 1) % 16 == 0:
        print(line); line = ""
print(line)
hot = sum(v[1] for v in chunks.values())
touched = sum(1 for v in chunks.values() if v[1] > 0)
print(f"instructions {len(recs)}, hot (executed by >= {frac:.0%} of the warps) {hot} = {hot * 16 / 1024:.1f} KB, "
      f"spread over {touched} chunks = {touched * 2} KB")
