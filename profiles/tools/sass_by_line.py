#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS export with `nvdisasm -g -c` line info and aggregate per CUDA source line.

usage: sass_by_line.py <ncu_source.csv> <nvdisasm.txt> <mangled-kernel-substring> [top]
Columns printed: executed warp instructions, stall samples, source line, text.
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    src_csv, dis, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    # nvdisasm: offset -> line
    off2line = {}
    cur = None
    inside = False
    for ln in open(dis):
        if ln.startswith(".text."):
            inside = kern in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            if "inlined at" not in ln:
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
        if m:
            off2line[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(src_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = None
    agg = defaultdict(lambda: [0, 0, 0])
    tot_i = tot_s = 0
    for r in rows[hdr_i + 1:]:
        if len(r) <= ii or not r[ia].startswith("0x"):
            continue
        a = int(r[ia], 16)
        if base is None:
            base = a
        line = off2line.get(a - base)
        n, s = int(r[ii] or 0), int(r[isamp] or 0)
        agg[line][0] += n; agg[line][1] += s; agg[line][2] += 1
        tot_i += n; tot_s += s
    srcs = {}
    print(f"total warp-instructions {tot_i}, samples {tot_s}")
    for line, (n, s, cnt) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if line:
            f = line[0]
            if f not in srcs:
                try:
                    import glob
                    p = glob.glob(f"/root/repo/**/{f}", recursive=True)
                    srcs[f] = open(p[0]).read().split("\n") if p else []
                except Exception:
                    srcs[f] = []
            if 0 < line[1] <= len(srcs[f]):
                text = srcs[f][line[1] - 1].strip()[:110]
        print(f"{n:9d} {100*n/tot_i:5.1f}%  samp {s:5d} {100*s/max(tot_s,1):5.1f}%  sass {cnt:4d}  {line}  {text}")


if __name__ == "__main__":
    main()
