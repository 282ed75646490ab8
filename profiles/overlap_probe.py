#!/usr/bin/env python
"""How much of the fp64 recheck hides behind the main-stream kernels: steps of FramePipeline.process on one B200 with the
recheck serialised (RS_NO_OVERLAP=1), pipelined across steps (join=False; the recheck runs beside the next step's 2-D FFT), pipelined with
the recheck held back until the next step's angle scan (RS_RECHECK_LATE=1), and with fewer resident CTAs of the scan.  CUDA-event time per 1000-frame step."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from radar_slam_b200 import FramePipeline, RadarConfig, synth
    cfg = RadarConfig(chirp_duration=256 / 10e6, num_chirps=128, num_antennas=8, search_resolution=1.0)
    pipe = FramePipeline(cfg)
    F = int(os.environ.get("FRAMES", "1000"))
    cube = synth.synth_cubes(cfg, F, seed=7, first_frame=0, device=pipe.device)
    vel = torch.empty((F, 8), dtype=torch.float64, device=pipe.device)
    variants = [("serial", {"RS_NO_OVERLAP": "1"}, True, F), ("pipelined", {}, False, F),
                ("pipelined, recheck beside the angle scan", {"RS_RECHECK_LATE": "1"}, False, F), ("no recheck", {"NOREC": "1"}, False, F)]
    for n in (7, 6, 5, 4):
        variants.append((f"pipelined, angles {n} CTAs/SM", {"RS_ANGLES_CTAS": str(n)}, False, F))
        variants.append((f"no recheck, angles {n} CTAs/SM", {"RS_ANGLES_CTAS": str(n), "NOREC": "1"}, False, F))
    for name, env, join, chunk in variants:
        for k in ("RS_NO_OVERLAP", "RS_RECHECK_LATE", "RS_ANGLES_CTAS"):
            os.environ.pop(k, None)
        os.environ.update({k: v for k, v in env.items() if k.startswith("RS_")})
        pipe.cfg.recheck = "NOREC" not in env
        for _ in range(3):
            pipe.process(cube, chunk_frames=chunk, vel_out=vel, join=join)
        pipe.join()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            pipe.process(cube, chunk_frames=chunk, vel_out=vel, join=join)
        pipe.join()
        e1.record()
        torch.cuda.synchronize()
        print(f"{name:28s} {e0.elapsed_time(e1) / n:.3f} ms/step", flush=True)


if __name__ == "__main__":
    main()
