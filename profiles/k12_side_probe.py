#!/usr/bin/env python
"""Probe: can a second, concurrently launched cluster kernel use the 16 SMs the 4-CTA-cluster persistent kernel strands
(33 clusters = 132 of 148 SMs)?  Main stream: ws kernel on frames [0, F - n_side); side stream: round 1's kernel
(RS_FUSED_NC = 2 or 4) on the last n_side frames.  CUDA-event time of the pair, per 1000 frames."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radar_slam_b200 import FramePipeline, RadarConfig, synth  # noqa: E402

F, A = 1000, 8
cfg = RadarConfig(chirp_duration=256 / 10e6, num_chirps=128, num_antennas=A)
pipe = FramePipeline(cfg)
cube = synth.synth_cubes(cfg, F, seed=7, first_frame=0, device=pipe.device)
out = torch.empty((F, 256, A, 128), dtype=torch.complex64, device=pipe.device)
side = torch.cuda.Stream()
main = torch.cuda.current_stream()


def run(n_side, nc, store, variant):
    def once():
        os.environ.update({"RS_K12": "ws", "RS_K12_STORE": store, "RS_K12_VARIANT": variant})
        fork = torch.cuda.Event()
        fork.record(main)
        pipe.range_doppler(cube[: F - n_side], out=out[: F - n_side])
        if n_side:
            os.environ.update({"RS_K12": "v1", "RS_FUSED_NC": nc})
            with torch.cuda.stream(side):
                side.wait_event(fork)
                pipe.range_doppler(cube[F - n_side:], out=out[F - n_side:])
                join = torch.cuda.Event()
                join.record(side)
            main.wait_event(join)
    for _ in range(3):
        once()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        once()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20


for store, variant in (("tma", "3"), ("direct", "4")):
    for nc in ("2", "4"):
        for n_side in (0, 40, 60, 80, 100):
            print(f"store={store} variant={variant} side: NC={nc} frames={n_side:3d}  ->  {run(n_side, nc, store, variant):.4f} ms", flush=True)
