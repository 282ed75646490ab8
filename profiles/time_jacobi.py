#!/usr/bin/env python
"""configs[3]-style stress of the general-covariance path: n Hermitian A x A covariances (4 snapshots each) through
rs_music_covariance (Jacobi eigendecomposition in registers + noise-subspace scan).  python profiles/time_jacobi.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from radar_slam_b200 import RadarConfig, FramePipeline

for A, n in ((8, 2_000_000), (16, 500_000)):
    pipe = FramePipeline(RadarConfig(num_antennas=A, search_resolution=1.0))
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((n, A, 4, 2), device="cuda", generator=g)
    s = torch.view_as_complex(x)
    cov = (s @ s.conj().transpose(1, 2)).to(torch.complex64).contiguous()
    pipe.music_covariance(cov, num_sources=2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = pipe.music_covariance(cov, num_sources=2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"A={A}: {n} covariances in {ms:.2f} ms = {n / ms / 1e3:.2f} M matrices/s (eigenvalues + 181-point spectrum + argmax)")
