import sys, numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from golden_util import CASE_NAMES, load_case, params_of, make_input
from oracle import radar_oracle as orc
from radar_slam_b200 import RadarConfig, FramePipeline
for name in CASE_NAMES:
    g, cfg = load_case(name); p = params_of(cfg); cube = make_input(cfg)
    pipe = FramePipeline(RadarConfig(chirp_duration=p.chirp_duration, num_chirps=p.num_chirps, num_antennas=p.num_antennas, window_type=p.window_type, dc_removal=p.dc_removal))
    rds = pipe.range_doppler(torch.from_numpy(cube[None]).cuda())[0].permute(2,0,1).cpu().numpy().astype(np.complex128)
    ref = orc.range_doppler_spectrum(cube.astype(np.complex128), p)
    rms = np.sqrt(np.mean(np.abs(ref)**2)); err = np.abs(rds-ref)
    print(name, "max_err/rms=%.3e rms_err/rms=%.3e max|X|/rms=%.1f" % (err.max()/rms, np.sqrt(np.mean(err**2))/rms, np.abs(ref).max()/rms))
