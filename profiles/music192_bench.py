import os, sys, torch, time
sys.path.insert(0, '/root/repo')
from radar_slam_b200 import FramePipeline, RadarConfig, synth
cfg = RadarConfig(chirp_duration=512 / 10e6, num_chirps=256, num_antennas=192, search_resolution=1.0, method="music")
pipe = FramePipeline(cfg)
cube = synth.synth_cubes(cfg, 4, seed=7, first_frame=0, device=pipe.device)
rds = pipe.range_doppler(cube); del cube
det = pipe.detect(rds)
for name, env in (("tcgen05", "1"), ("cuda-core", "0")):
    os.environ["RS_MUSIC_TC"] = env
    pipe.angles(rds, det); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pipe.angles(rds, det); e1.record(); torch.cuda.synchronize()
    print(name, "ms per 4 frames of 512x256x192:", round(e0.elapsed_time(e1), 3), flush=True)
