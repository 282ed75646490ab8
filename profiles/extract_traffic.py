#!/usr/bin/env python
"""Per-frame DRAM traffic of every hot-path stage from an `ncu --set full` capture (bench.py reports it as
roofline.traffic):   python profiles/extract_traffic.py gpurun_out/ncu_<tag> <frames per launch> profiles/<tag>_traffic.json"""
import csv
import io
import json
import subprocess
import sys

STAGE = {"recheck_detect": "rs_recheck_detections_f64", "range_fft": "rs_range_fft",
         "doppler_fft": "rs_doppler_fft", "detect_kernel": "rs_detect", "angles_": "rs_angles",
         "velocity_from": "rs_velocity_from_partials"}
RECHECK_PARTS = ("recheck_angles", "recheck_snapshots", "recheck_finish")
K12_PARTS = ("fft2d_ws", "fft2d_cluster")        # one rs_range_doppler_fft call = persistent cluster kernel + side kernel
FUSED_PARTS = ("compact_masks", "detect_a8")      # rs_range_doppler_detect = the two above (detection fused) + compaction + side frames' detect
MULT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    d, frames, out = sys.argv[1], float(sys.argv[2]), sys.argv[3]
    raw = subprocess.run(["ncu", "-i", f"{d}/prof.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki, rd, wr, gi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("launch__grid_size")
    num = lambda x: float(x.replace(",", ""))
    # the capture window may run past the benchmark step into smaller launch sets (the end-to-end chunks): per kernel keep
    # only the launches that executed about as many instructions as the largest one
    ii = hdr.index("smsp__inst_executed.sum")
    biggest = {}
    for r in rows[2:]:
        biggest[r[ki]] = max(biggest.get(r[ki], 0.0), num(r[ii]))
    acc, parts, k12 = {}, {}, {}
    for r in rows[2:]:
        if num(r[ii]) < 0.6 * biggest[r[ki]]:
            continue
        b = num(r[rd]) * MULT[units[rd]] + num(r[wr]) * MULT[units[wr]]
        hit12 = [k for k in K12_PARTS + FUSED_PARTS if k in r[ki]]
        if hit12:
            k12.setdefault(hit12[0], []).append(b)
            continue
        hit = [k for k in RECHECK_PARTS if k in r[ki]]
        if hit:
            nm = r[ki]
            nm = nm[: nm.index(">(") + 1] if ">(" in nm else nm.split("(")[0]
            parts.setdefault(nm, []).append(b)                       # template instantiations are separate launches of one call
            continue
        for k, v in STAGE.items():
            if k in r[ki]:
                acc.setdefault(v, []).append(b)
                break
    per_frame = {k: sum(v) / len(v) / frames for k, v in acc.items()}
    if k12:
        fused = "compact_masks" in k12
        if not fused and "detect_a8" in k12:
            per_frame["rs_detect"] = sum(k12["detect_a8"]) / len(k12.pop("detect_a8")) / frames
        per_frame["rs_range_doppler_detect" if fused else "rs_range_doppler_fft"] = sum(sum(v) / len(v) for v in k12.values()) / frames
        if fused:
            per_frame["parts_of_rs_range_doppler_detect"] = {k: sum(v) / len(v) / frames for k, v in k12.items()}
    if parts:
        per_frame["rs_recheck_angles_f64"] = sum(sum(v) / len(v) for v in parts.values()) / frames
    json.dump({"source": f"{d}/prof.ncu-rep (ncu --set full --clock-control none): dram__bytes_read.sum + dram__bytes_write.sum "
                         f"per launch of {frames:g} frames, divided by the frames",
               "dram_bytes_per_frame": per_frame}, open(out, "w"), indent=1)
    print(json.dumps(per_frame, indent=1))


if __name__ == "__main__":
    main()
