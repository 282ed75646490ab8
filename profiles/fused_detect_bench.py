#!/usr/bin/env python
"""2-D FFT + detection: the fused call (rs_range_doppler_detect) against the two stages, CUDA-event times on one B200.

    python profiles/fused_detect_bench.py [--frames 1000] [--antennas 8] [--reps 10] [--sides 0,30,50,60]

Prints one JSON line per variant: ms per launch set and the algorithmic GB/s of the 2-D transform (16 B per cell).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--antennas", type=int, default=8)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--sides", default="0,30,50,60")
    args = ap.parse_args()
    from radar_slam_b200 import FramePipeline, RadarConfig, synth, _lib

    cfg = RadarConfig(chirp_duration=256 / 10e6, num_chirps=128, num_antennas=args.antennas, search_resolution=1.0)
    pipe = FramePipeline(cfg)
    F, A = args.frames, args.antennas
    cube = synth.synth_cubes(cfg, F, seed=7, first_frame=0, device=pipe.device)
    rds = torch.empty((F, 256, A, 128), dtype=torch.complex64, device=pipe.device)
    peak = 6446.3
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk)).get("hbm_gbs", peak)
    nbytes = 16 * cube.numel()

    def report(name, ms, **kw):
        print(json.dumps(dict(variant=name, ms=round(ms, 4), fft_alg_gbs=round(nbytes / ms / 1e6, 1),
                              frac_hbm=round(nbytes / ms / 1e6 / peak, 3), **kw)), flush=True)

    for k in ("RS_K12_SIDE", "RS_FUSED_DETECT", "RS_SPLIT_DETECT"):
        os.environ.pop(k, None)
    t_fft = timed(lambda: pipe.range_doppler(cube, out=rds), args.reps)
    report("fft only (default side share)", t_fft)
    det = pipe.detect(rds, workspace="0")
    t_det = timed(lambda: pipe.detect(rds, workspace="0"), args.reps)
    report("detect only", t_det, detections_per_frame=float(det.per_frame_counts().float().mean()))
    report("fft + detect, two stages", timed(lambda: (pipe.range_doppler(cube, out=rds), pipe.detect(rds, workspace="0")), args.reps))
    for side in args.sides.split(","):
        os.environ["RS_K12_SIDE"] = side
        ms = timed(lambda: pipe.range_doppler_detect(cube, out=rds, workspace="0"), args.reps)
        report(f"fused, side {side} permille", ms)
        ms = timed(lambda: pipe.range_doppler_detect(cube, out=rds, workspace="0", defer_power=True), args.reps)
        report(f"fused without det_power, side {side} permille", ms)
        os.environ["RS_FD_NO_COMPACT"] = "1"
        ms = timed(lambda: pipe.range_doppler_detect(cube, out=rds, workspace="0", defer_power=True), args.reps)
        report(f"fused kernel only (no compaction), side {side} permille", ms)
        os.environ.pop("RS_FD_NO_COMPACT")


if __name__ == "__main__":
    main()
