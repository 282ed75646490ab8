#!/bin/bash
# What does the detection cost inside the fused 2-D FFT kernel?  Timing ablation (results are WRONG with RS_FD_DBG != 0):
# 1 = no walk, 2 = no neighbour barriers / halo wait, 4 = no power stores.  One B200, 1000 frames, no side kernel.
for d in 0 1 2 3 4 7; do echo "RS_FD_DBG=$d"; RS_FD_DBG=$d RS_FD_NO_COMPACT=1 python - <<'PY'
import os,sys,torch
sys.path.insert(0,'.')
from radar_slam_b200 import FramePipeline, RadarConfig, synth
cfg = RadarConfig(chirp_duration=256/10e6, num_chirps=128, num_antennas=8, search_resolution=1.0)
pipe = FramePipeline(cfg); F=1000
cube = synth.synth_cubes(cfg, F, seed=7, first_frame=0, device=pipe.device)
rds = torch.empty((F,256,8,128), dtype=torch.complex64, device=pipe.device)
os.environ["RS_K12_SIDE"]="0"
f=lambda: pipe.range_doppler_detect(cube, out=rds, workspace="0", defer_power=True)
for _ in range(3): f()
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): f()
e1.record(); torch.cuda.synchronize()
print("  fused kernel: %.4f ms"%(e0.elapsed_time(e1)/10))
g=lambda: pipe.range_doppler(cube, out=rds)
for _ in range(3): g()
e0.record()
for _ in range(10): g()
e1.record(); torch.cuda.synchronize()
print("  fft only (side 0): %.4f ms"%(e0.elapsed_time(e1)/10))
PY
done
