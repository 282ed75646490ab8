#!/usr/bin/env python
"""Print a one-line digest of a bench.py JSON line read from stdin (tuning helper)."""
import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    st = " ".join(f"{s['kernel'].replace('rs_','')}={s['ms_per_step']:.2f}" for s in d["roofline"]["stages"])
    print(f"{sys.argv[1] if len(sys.argv) > 1 else ''} value={d['value']:.0f} e2e={d['e2e']['value']:.0f} ms/step={d['ms_per_step']:.2f} | {st}")
