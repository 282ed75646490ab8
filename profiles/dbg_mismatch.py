import sys, os, numpy as np, torch
sys.path.insert(0, "/root/repo")
from oracle import radar_oracle as orc
from radar_slam_b200 import RadarConfig, FramePipeline, synth
F, S, C, A = 8, 256, 128, 8
cfg = RadarConfig(chirp_duration=S / 10e6, num_chirps=C, num_antennas=A, search_resolution=1.0, method="music")
pipe = FramePipeline(cfg)
cube = synth.synth_cubes(cfg, 500, seed=77)[:F].contiguous()
p = orc.RadarParams(chirp_duration=S / 10e6, num_chirps=C, num_antennas=A)
grid = orc.azimuth_grid((-90, 90), 1.0)
steer = orc.steering_matrix(grid, p.antenna_positions, p.lambda_c)
tot = 0
for f in range(F):
    vel, rds, det = pipe.process(cube[f:f + 1], keep=True)
    frame = cube[f].cpu().numpy()
    ref = orc.range_doppler_spectrum(frame.astype(np.complex128), p)
    pk = orc.extract_peaks(ref, p, threshold_db=-20.0)
    d = det.frame(0)
    sigs = orc.spatial_signatures(ref, pk["range_bin"], pk["doppler_bin"])
    spec = orc.beamforming_spectra(sigs, steer)
    den = np.abs(A - spec)
    with np.errstate(divide="ignore"):
        mus = np.where(den > 1e-12, 1.0 / den, 0.0)
    idx = np.argmax(mus, axis=1)
    bad = np.nonzero(d["aidx"] != idx)[0]
    tot += len(bad)
    for b in bad[:6]:
        g0, g1 = int(d["aidx"][b]), int(idx[b])
        srt = np.sort(spec[b])[::-1]
        print(f"frame {f} det {b} key {d['key'][b]:08x} flags {d['flags'][b]:02x} gpu {g0} oracle {g1} P[gpu] {spec[b,g0]:.9f} P[or] {spec[b,g1]:.9f} relgap {(spec[b,g1]-spec[b,g0])/spec[b,g1]:.3e} top2gap {(srt[0]-srt[1])/srt[0]:.3e}")
    st = pipe._ws["recheck_ang_stats"].cpu().numpy()
    print("frame", f, "mismatches", len(bad), "of", len(idx), "stats", st)
print("total mismatches", tot, "MMA", os.environ.get("RS_ANGLES_MMA"))
