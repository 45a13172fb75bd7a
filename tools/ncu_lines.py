#!/usr/bin/env python3
"""Aggregate an ncu report's warp-stall samples and executed instructions by CUDA source line.
usage: tools/ncu_lines.py report.ncu-rep [top_n]"""
import collections, csv, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ii, wi = hdr.index("Instructions Executed"), 4
cur, agg, fname = None, collections.OrderedDict(), ""
for r in rows[hi + 1:]:
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]
        continue
    if r[0]:
        try:
            cur = (fname, int(r[0]), r[1][:100])
        except ValueError:
            continue
        agg.setdefault(cur, [0, 0])
    elif cur and len(r) > ii and r[2] not in ("...", "-", ""):
        try:
            agg[cur][0] += int(r[ii] or 0)
            agg[cur][1] += int(r[wi] or 0)
        except ValueError:
            pass
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
print(f"instructions {ti}  stall samples {ts}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100 * v[1] / ts:5.1f}% samples {100 * v[0] / ti:5.1f}% inst  {k[0][:14]}:{k[1]:>4} {k[2]}")
