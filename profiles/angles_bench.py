#!/usr/bin/env python
"""Angle-scan variants side by side on one B200: mma.sync (default) against the tcgen05 / TMEM scan (RS_ANGLES_TC=1).
Same RDS, same detection lists; reports CUDA-event time per pass and how many cells differ in grid index / flags."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--antennas", type=int, default=8)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--grid-res", type=float, default=1.0)
    args = ap.parse_args()
    from radar_slam_b200 import FramePipeline, RadarConfig, synth
    cfg = RadarConfig(chirp_duration=256 / 10e6, num_chirps=128, num_antennas=args.antennas, search_resolution=args.grid_res)
    pipe = FramePipeline(cfg)
    cube = synth.synth_cubes(cfg, args.frames, seed=7, first_frame=0, device=pipe.device)
    rds = pipe.range_doppler(cube)
    del cube
    det = pipe.detect(rds)
    n = det.F * det.ntiles * det.seg_cap
    ref = None
    for name, env in (("mma.sync", {"RS_ANGLES_TC": "0"}), ("tcgen05", {"RS_ANGLES_TC": "1"}), ("cuda-core", {"RS_ANGLES_TC": "0", "RS_ANGLES_MMA": "0"})):
        os.environ.pop("RS_ANGLES_MMA", None)
        os.environ.update(env)
        flags0 = det.flags.clone()
        for _ in range(2):
            det.flags.copy_(flags0)
            pipe.angles(rds, det)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ms = 0.0
        for _ in range(args.reps):
            det.flags.copy_(flags0)
            e0.record()
            pipe.angles(rds, det)
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1) / args.reps
        m = det.valid_mask().reshape(-1)
        aidx = det.aidx[:n][m].clone()
        fl = det.flags[:n][m].clone()
        part = det.ls_partials.clone()
        det.flags.copy_(flags0)
        if ref is None:
            ref = (aidx, fl, part)
        out = {"variant": name, "ms": round(ms, 4), "detections": int(m.sum()),
               "aidx_differs": int((aidx != ref[0]).sum()), "flags_differ": int((fl != ref[1]).sum()),
               "aidx_differs_unflagged": int(((aidx != ref[0]) & ((fl & 5) == 0) & ((ref[1] & 5) == 0)).sum()),
               "ls_partials_max_abs_diff": float((part - ref[2]).abs().max())}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
