#!/usr/bin/env python
"""Digest of `ncu -i X.ncu-rep --page source --csv` (SASS view): instruction mix, stall samples per code region
(regions = runs of SASS lines with the same execution-count class) and the top stall sites.
    python profiles/sass_regions.py src.csv [raw.csv]"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]; ci = h.index('Instructions Executed'); si = h.index('Source'); sm = h.index('# Samples')
data = []
for k, r in enumerate(rows[2:]):
    try: data.append((k, int(r[ci]), int(r[sm]), r[si]))
    except Exception: pass
tot = sum(d[2] for d in data); ninst = sum(d[1] for d in data)
print(f"instructions {ninst/1e6:.1f} M, samples {tot}")
by = collections.Counter(); sb = collections.Counter()
for k, n, s, src in data:
    t = src.split(); op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    by[op] += n; sb[op] += s
print("mix: " + ", ".join(f"{k} {100*v/ninst:.1f}%/{100*sb[k]/tot:.1f}%s" for k, v in by.most_common(14)))
counts = sorted({n for _, n, _, _ in data if n > 0})
top = max(counts)
def cls(n):
    return 'once' if n < top / 20 else ('inner' if n > top * 0.7 else 'tile')
regions = []; cur = None
for k, n, s, src in data:
    c = cls(n)
    if cur is None or cur[0] != c:
        cur = [c, k, k, 0, 0, 0]; regions.append(cur)
    cur[2] = k; cur[3] += n; cur[4] += s; cur[5] += 1
for c in regions:
    if c[4] > 0.01 * tot or c[3] > 0.02 * ninst:
        print(f"{c[0]:5s} sass#{c[1]:5d}-{c[2]:5d} lines={c[5]:4d} exec={100*c[3]/ninst:5.1f}% samples={100*c[4]/tot:5.1f}%")
for k, n, s, src in sorted(data, key=lambda d: -d[2])[:16]:
    print(f"{100*s/tot:5.1f}% exec={n:8d} #{k:5d} {src[:100]}")
if len(sys.argv) > 2:
    rows = list(csv.reader(open(sys.argv[2]))); h = rows[0]; r = rows[2]
    for key in ('gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
                'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum'):
        if key in h: print(key, r[h.index(key)], rows[1][h.index(key)])
    for i, k in enumerate(h):
        if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and float(r[i] or 0) > 0.15:
            print('  stall', k.split('issue_stalled_')[1].split('_per_')[0], r[i])
