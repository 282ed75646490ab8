#!/usr/bin/env python
"""K12 (fused 2-D FFT) variants side by side on one B200: correctness against round 1's kernel and CUDA-event times.

    python profiles/k12_bench.py [--frames 1000] [--antennas 8] [--reps 10] [--variants v1,ws,ws-tma]

Each variant transforms the same device-resident cube (F x A x 128 x 256 complex64, larger than L2) `reps` times;
the line reports the mean launch time, algorithmic GB/s (16 B per cell) and the fraction of the measured HBM peak.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = {
    "v1": {"RS_K12": "v1"},
    "ws": {"RS_K12": "ws"},                                           # library default
    "ws-tma": {"RS_K12": "ws", "RS_K12_STORE": "tma"},
    "split": {"RS_FUSED_FFT": "0"},
}
for _v in range(8):                                                    # wsN / wsN-tma: RS_K12_VARIANT = N (rs_fft2d_ws.cu)
    VARIANTS[f"ws{_v}"] = {"RS_K12": "ws", "RS_K12_VARIANT": str(_v), "RS_K12_STORE": "direct"}
    VARIANTS[f"ws{_v}-tma"] = {"RS_K12": "ws", "RS_K12_VARIANT": str(_v), "RS_K12_STORE": "tma"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--antennas", type=int, default=8)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--variants", default="v1,ws,ws-tma")
    ap.add_argument("--clusters", default="")
    args = ap.parse_args()
    from radar_slam_b200 import FramePipeline, RadarConfig, synth, _lib

    cfg = RadarConfig(chirp_duration=256 / 10e6, num_chirps=128, num_antennas=args.antennas)
    pipe = FramePipeline(cfg)
    cube = synth.synth_cubes(cfg, args.frames, seed=7, first_frame=0, device=pipe.device)
    peak = 6446.3
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk)).get("hbm_gbs", peak)
    print("max co-resident clusters of the ws kernel:", _lib.load().rs_fft2d_ws_max_clusters(), flush=True)
    ref = None
    nbytes = 16 * cube.numel()
    for name in args.variants.split(","):
        for k in ("RS_K12", "RS_K12_STORE", "RS_K12_VARIANT", "RS_K12_SIDE", "RS_FUSED_FFT", "RS_FUSED_NC", "RS_K12_CLUSTERS"):
            os.environ.pop(k, None)
        base, _, ncl = name.partition("@")
        base, _, side = base.partition("+")                           # ws3-tma+60: 60 permille of the frames to the side kernel
        os.environ.update(VARIANTS[base])
        os.environ["RS_K12_SIDE"] = side or "0"
        if ncl:
            os.environ["RS_K12_CLUSTERS"] = ncl
        out = torch.empty((args.frames, 256, args.antennas, 128), dtype=torch.complex64, device=pipe.device)
        for _ in range(3):
            pipe.range_doppler(cube, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            pipe.range_doppler(cube, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        if ref is None:
            ref = out.clone()
            err = 0.0
        else:
            err = float((out - ref).abs().max() / ref.abs().max())
        print(json.dumps({"variant": name, "ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1),
                          "frac_hbm": round(nbytes / ms / 1e6 / peak, 4), "max_rel_diff_vs_first": err}), flush=True)


if __name__ == "__main__":
    main()
