#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) and a launch-list csv into a small markdown summary.

    python profiles/summarize_ncu.py gpurun_out/ncu_<tag> profiles/<tag>_summary.md
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = [
    ("gpu__time_duration.sum", "dur_us", 1e-3),
    ("dram__bytes_read.sum", "dram_rd_MB", 1e-6),
    ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct", 1),
    ("lts__t_bytes.sum", "l2_MB", 1e-6),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__grid_size", "grid", 1),
    ("launch__block_size", "block", 1),
    ("smsp__inst_executed.sum", "inst_M", 1e-6),
    ("sm__inst_executed_pipe_fma.sum", "fma_inst_M", 1e-6),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pct", 1),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct", 1),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts_M", 1e-6),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts_M", 1e-6),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall_long_sb", 1),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "stall_short_sb", 1),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall_barrier", 1),
    ("smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "stall_mio", 1),
    ("smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "stall_lg", 1),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "stall_math", 1),
    ("smsp__average_warp_latency_issue_stalled_wait.ratio", "stall_wait", 1),
]


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main():
    d, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", f"{d}/prof.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    units = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    lines = ["# ncu summary: " + d, "", "## full capture (ncu --set full --clock-control none), one row per profiled launch", ""]
    names = [k for k, _, _ in KEYS if k in col]
    short = {k: s for k, s, _ in KEYS}
    scale = {k: f for k, _, f in KEYS}
    lines.append("| kernel | " + " | ".join(short[k] for k in names) + " |")
    lines.append("|---|" + "---|" * len(names))
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        kn = r[col["Kernel Name"]].split("(")[0][-40:]
        vals = []
        for k in names:
            v = num(r[col[k]])
            u = units[col[k]]
            if v is None:
                vals.append("-")
                continue
            if k == "gpu__time_duration.sum":
                v = v / 1e3 if u in ("ns", "nsecond") else v * (1e3 if u in ("ms", "msecond") else 1)
                vals.append(f"{v:.1f}")
                continue
            if k.startswith("dram__bytes") or k.startswith("lts__t_bytes"):
                mult = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1e-6)
                vals.append(f"{v * mult:.2f}")
                continue
            vals.append(f"{v * scale[k]:.3g}")
        lines.append(f"| {kn} | " + " | ".join(vals) + " |")
    # launch list
    try:
        txt = open(f"{d}/launches.csv").read()
        start = txt.index('"ID"')
        lr = list(csv.DictReader(io.StringIO(txt[start:])))
        agg = defaultdict(list)
        for r in lr:
            if r.get("Metric Name") == "gpu__time_duration.sum":
                v = num(r["Metric Value"])
                u = r.get("Metric Unit", "ns")
                v = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
                agg[r["Kernel Name"].split("(")[0][-40:]].append(v)
        tot = sum(sum(v) for v in agg.values())
        lines += ["", "## launch list (ncu --metrics gpu__time_duration.sum; cold-cache, serialised: compare shares)", "",
                  "| kernel | launches | mean us | total us | share |", "|---|---|---|---|---|"]
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            lines.append(f"| {k} | {len(v)} | {sum(v) / len(v):.1f} | {sum(v):.1f} | {sum(v) / tot:.3f} |")
    except Exception as e:  # noqa
        lines.append(f"(launch list not parsed: {e})")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
