#!/usr/bin/env python
"""Static count of the Blackwell / Hopper-class SASS mnemonics per kernel of the shipped sm_100a binary:
    python profiles/sass_mnemonics.py > profiles/r02_sass_mnemonics.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "radar_slam_b200", "lib", "libradarslam_b200.so")
COLS = ["UTMALDG", "UBLKCP", "STAS", "SYNCS", "UCGABAR_ARV", "UTCHMMA", "LDTM", "HMMA", "FADD2", "FFMA2", "FMNMX3"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    counts, cur, it = {}, None, iter(names)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            full = next(it)
            short = re.sub(r"\(int\)|\(bool\)", "", full.replace("(anonymous namespace)::", "").replace("void ", ""))
            short = short.split("(")[0]
            cur = counts.setdefault(short, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            op = m.group(1)
            for c in COLS:
                if op == c:
                    cur[c] += 1
    print("# Blackwell / Hopper-class SASS mnemonics in the shipped sm_100a binary (round 2)\n")
    print("`python profiles/sass_mnemonics.py` (`cuobjdump -sass radar_slam_b200/lib/libradarslam_b200.so`), instructions counted per kernel (static counts).")
    print("UTMALDG = TMA tensor load (`cp.async.bulk.tensor`), UBLKCP = bulk (TMA) store (`cp.async.bulk.global.shared::cta`), STAS = `st.async` into a peer CTA's")
    print("shared memory completing on its mbarrier, SYNCS = mbarrier operations, UCGABAR = cluster barrier, UTCHMMA = `tcgen05.mma` (UMMA),")
    print("LDTM = `tcgen05.ld` (TMEM -> registers), HMMA = legacy `mma.sync`, FADD2 / FFMA2 = packed f32x2 arithmetic, FMNMX3 = 3-input min/max.\n")
    print("| kernel | " + " | ".join(COLS) + " |")
    print("|---|" + "---|" * len(COLS))
    for k in sorted(counts):
        if any(counts[k][c] for c in COLS):
            print(f"| `{k}` | " + " | ".join(str(counts[k][c]) if counts[k][c] else "" for c in COLS) + " |")


if __name__ == "__main__":
    main()
