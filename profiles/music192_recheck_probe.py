import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from oracle import radar_oracle as orc
from radar_slam_b200 import RadarConfig, FramePipeline
S, C, A = 64, 128, 192
p = orc.RadarParams(chirp_duration=S / 10e6, num_chirps=C, num_antennas=A)
scene = np.array([(3.0, 0.3, 10.0, 0.0), (6.0, -0.6, 0.0, 0.0)])
np.random.seed(1)
cube = np.stack([orc.synthesize_frame(p, scene) for _ in range(2)]).astype(np.complex64)
grid = orc.azimuth_grid((-90, 90), 1.0)
steer = orc.steering_matrix(grid, p.antenna_positions, p.lambda_c)
cfg = RadarConfig(chirp_duration=S / 10e6, num_chirps=C, num_antennas=A, search_resolution=1.0, method="music", recheck=True)
pipe = FramePipeline(cfg)
vel, rds, det = pipe.process(torch.from_numpy(cube).cuda(), keep=True)
torch.cuda.synchronize()
print("status", pipe.status())
for f in range(2):
    ref = orc.range_doppler_spectrum(cube[f].astype(np.complex128), p)
    d = det.frame(f)
    sub = np.arange(0, len(d["key"]), max(1, len(d["key"]) // 3000))
    sigs = orc.spatial_signatures(ref, d["range_bin"][sub], d["doppler_bin"][sub])
    idx, _ = orc.argmax_angles(orc.music_spectra(sigs, steer), grid)
    bad = d["aidx"][sub] != idx
    print(f, "detections", len(d["key"]), "sample", len(sub), "mismatch", int(bad.sum()), "flagged among them", int((d["flags"][sub][bad] & 5 != 0).sum()),
          "TIE share", float(np.mean(d["flags"] & 1 != 0)), "FIXED share", float(np.mean(d["flags"] & 8 != 0)))
