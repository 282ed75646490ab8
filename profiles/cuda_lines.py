#!/usr/bin/env python
"""Stall samples and executed instructions per CUDA source line from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass -k regex:<kernel> -c 1 > lines.csv`:
    python profiles/cuda_lines.py lines.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, out = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[0] == "":
        continue
    try:
        out.append((int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")]), cur_file, int(r[0]), r[1].strip()))
    except ValueError:
        pass
ts = sum(o[0] for o in out) or 1; ti = sum(o[1] for o in out) or 1
print(f"samples {ts}, instructions {ti / 1e6:.1f} M")
for s, n, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{100 * s / ts:5.1f}% samples {100 * n / ti:5.1f}% instr  {f}:{ln}  {src[:110]}")
