#!/usr/bin/env python
"""Timeline of the stage calls of pipelined FramePipeline.process steps (CUDA events on both streams, relative to one
origin): shows which main-stream kernels the side-stream recheck actually runs beside."""
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
    for _ in range(3):
        pipe.process(cube, chunk_frames=F, vel_out=vel, join=False)
    pipe.join()
    torch.cuda.synchronize()
    origin = torch.cuda.Event(enable_timing=True)
    origin.record()
    pipe.profile = []
    for _ in range(3):
        pipe.process(cube, chunk_frames=F, vel_out=vel, join=False)
    pipe.join()
    torch.cuda.synchronize()
    for name, a, b in pipe.profile:
        print(f"{name:28s} {origin.elapsed_time(a):8.3f} -> {origin.elapsed_time(b):8.3f}  ({a.elapsed_time(b):.3f} ms)")


if __name__ == "__main__":
    main()
