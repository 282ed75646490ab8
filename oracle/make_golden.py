#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- write tests/golden/*.npz from the LIVE reference classes.

Run in the build container (needs /root/reference):  python oracle/make_golden.py
The reference cannot travel to the GPU box, so its outputs on seeded inputs are committed as
small fixtures.  Inputs are NOT stored: they are regenerated from (params, seed, scatterers) by
oracle.radar_oracle.synthesize_frame, which is bit-identical to the reference simulator
(tests/test_oracle_vs_reference.py); a checksum of the cube guards that.

"Identical inputs" (BASELINE.json): the cube is generated in fp64, rounded to complex64 -- the
dtype the CUDA path consumes -- and that rounded cube (cast back to complex128) is what the
reference sees.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import radar_oracle as orc   # noqa: E402
from oracle import ref_import            # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

SCENE = np.array([(8.0, np.radians(0.0), -10.0, 0.0), (12.0, np.radians(30.0), -8.0, 0.0),
                  (16.0, np.radians(-20.0), -6.0, 0.0), (20.0, np.radians(10.0), -3.0, 0.0),
                  (25.0, np.radians(-40.0), 0.0, 0.0)])

CASES = {
    # name: (S, C, A, window, dc, seed, threshold_db, grid_res, n_angle_subsample, n_velocity, noise_power)
    "c1_default": dict(S=256, C=128, A=8, window="hann", dc=True, seed=1000, thr=-20.0, res=0.5, n_sub=400, n_vel=40, noise=0.01),
    "ref_default_400x64": dict(S=400, C=64, A=8, window="hann", dc=True, seed=1001, thr=-20.0, res=0.5, n_sub=300, n_vel=0, noise=0.01),
    "c4_sparse": dict(S=256, C=128, A=8, window="hann", dc=True, seed=1002, thr=30.5, res=1.0, n_sub=0, n_vel=40, noise=0.01),
    "small_hamming_nodc": dict(S=64, C=32, A=4, window="hamming", dc=False, seed=1003, thr=10.0, res=1.0, n_sub=0, n_vel=12, noise=0.01),
    "a16_blackman": dict(S=256, C=128, A=16, window="blackman", dc=True, seed=1004, thr=29.0, res=2.0, n_sub=0, n_vel=0, noise=0.01),
    "lownoise_400x32": dict(S=400, C=32, A=8, window="hann", dc=True, seed=1005, thr=-30.0, res=0.5, n_sub=300, n_vel=0, noise=1e-7),
}


def params_of(c) -> orc.RadarParams:
    return orc.RadarParams(chirp_duration=c["S"] / 10e6, num_chirps=c["C"], num_antennas=c["A"],
                           window_type=c["window"], dc_removal=c["dc"], noise_power=c["noise"])


def make_input(c) -> np.ndarray:
    np.random.seed(c["seed"])
    cube = orc.synthesize_frame(params_of(c), SCENE)
    return cube.astype(np.complex64)


def run_case(name, c, ref):
    p = params_of(c)
    cube64 = make_input(c)
    cube = cube64.astype(np.complex128)
    pre = ref.SignalPreprocessor(fc=p.fc, bandwidth=p.bandwidth, chirp_duration=p.chirp_duration, pri=p.pri,
                                 num_chirps=p.num_chirps, sampling_rate=p.sampling_rate,
                                 window_type=p.window_type, dc_removal=p.dc_removal)
    assert pre.samples_per_chirp == c["S"]
    rds = pre.generate_range_doppler_spectrum(cube)
    pi = pre.extract_range_doppler_peaks(rds, threshold_db=c["thr"])
    peaks = pi["peaks"]
    D = len(peaks)
    out = {
        "cube_checksum": np.array([cube64.real.astype(np.float64).sum(), cube64.imag.astype(np.float64).sum(),
                                   np.abs(cube64.astype(np.complex128)).sum()]),
        "rds_abs_max": np.abs(rds).max(),
        "rds_abs_sum": np.abs(rds).sum(),
        "pk_antenna": np.array([q["antenna"] for q in peaks], dtype=np.uint8),
        "pk_range_bin": np.array([q["range_bin"] for q in peaks], dtype=np.uint16),
        "pk_doppler_bin": np.array([q["doppler_bin"] for q in peaks], dtype=np.uint16),
        "range_bins_m": pi["range_bins_m"], "doppler_bins_hz": pi["doppler_bins_hz"],
    }
    rs = np.random.RandomState(c["seed"] + 77)
    if rds.size <= 1 << 14:
        out["rds_full"] = rds
    n_cells = min(4096, rds.size)
    flat = rs.choice(rds.size, n_cells, replace=False)
    out["rds_sample_idx"] = flat.astype(np.int64)
    out["rds_sample_val"] = rds.reshape(-1)[flat]
    sub_pw = np.arange(D) if D <= 4096 else np.sort(rs.choice(D, 4096, replace=False))
    out["pk_power_idx"] = sub_pw.astype(np.int64)
    out["pk_power_db"] = np.array([peaks[i]["power_db"] for i in sub_pw])

    # ---- angles on all peaks or a stated subsample
    sub = np.arange(D) if (c["n_sub"] == 0 or D <= c["n_sub"]) else np.sort(rs.choice(D, c["n_sub"], replace=False))
    est = ref.AngleEstimator(fc=p.fc, num_antennas=p.num_antennas, search_resolution=c["res"])
    grid = est.azimuth_grid
    sub_info = {"peaks": [peaks[i] for i in sub]}
    t0 = time.time()
    tm = est.process_targets(rds, sub_info, "music")
    te = est.process_targets(rds, sub_info, "esprit")
    tb = est.process_targets(rds, sub_info, "beamforming")
    assert len(tm) == len(sub)
    out["ang_sub"] = sub.astype(np.int64)
    out["grid_deg"] = grid
    out["music_deg"] = np.array([t["azimuth_deg"] for t in tm])
    out["esprit_deg"] = np.array([t["azimuth_deg"] for t in te])
    out["beam_deg"] = np.array([t["azimuth_deg"] for t in tb])
    out["sig_first8"] = np.array([t["spatial_signature"] for t in tm[:8]])
    out["music_spec_first2"] = np.array([t["spectrum"] for t in tm[:2]])
    # top-2 relative gap of the MUSIC spectrum (tie diagnostics for the tolerance statement)
    gaps = []
    for t in tm:
        s = np.sort(t["spectrum"])[::-1]
        gaps.append((s[0] - s[1]) / s[0] if s[0] > 0 else 0.0)
    out["music_top2_gap"] = np.array(gaps)
    print(f"  {name}: D={D} angles on {len(sub)} peaks in {time.time() - t0:.1f}s")

    # ---- velocity: reference DE on the n_vel strongest of the angle subsample (full-D DE takes hours)
    if c["n_vel"]:
        order = np.argsort([-t["power_db"] for t in tm], kind="stable")[: c["n_vel"]]
        tv = [tm[i] for i in order]
        lam = 3e8 / 77e9
        solver = ref.VelocitySolver(fc=p.fc, lambda_c=lam, num_antennas=p.num_antennas)
        t0 = time.time()
        res = solver.solve_velocity(rds, tv, dt=0.1)
        out["vel_sel"] = np.array([sub[i] for i in order], dtype=np.int64)
        out["vel_success"] = np.array(bool(res["success"]))
        if res["success"]:
            out["vel_velocity"] = np.asarray(res["velocity"])
            out["vel_cost"] = np.array(res["cost"])
            out["vel_observed"] = np.asarray(res["observed_phases"])
        print(f"     velocity (DE, N={len(tv)}) in {time.time() - t0:.1f}s -> {res.get('velocity')}")

    out["meta"] = np.array(repr(c))
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **out)


def run_robust(ref):
    """Three consecutive frames through RobustAngleEstimator (stateful smoothing)."""
    c = dict(S=256, C=64, A=8, window="hann", dc=True, seed=2000, thr=20.0, noise=0.01)
    p = params_of(c)
    pre = ref.SignalPreprocessor(fc=p.fc, bandwidth=p.bandwidth, chirp_duration=p.chirp_duration, pri=p.pri,
                                 num_chirps=p.num_chirps, sampling_rate=p.sampling_rate)
    rob = ref.RobustAngleEstimator(fc=p.fc, num_antennas=8, max_targets=50)
    out = {"meta": np.array(repr(c))}
    for k in range(3):
        cc = dict(c, seed=c["seed"] + k)
        cube = make_input(cc).astype(np.complex128)
        rds = pre.generate_range_doppler_spectrum(cube)
        pi = pre.extract_range_doppler_peaks(rds, threshold_db=c["thr"])
        tg = rob.process_targets_robust(rds, pi, frame_timestamp=float(k))
        out[f"f{k}_range_bin"] = np.array([t["range_bin"] for t in tg], dtype=np.int32)
        out[f"f{k}_doppler_bin"] = np.array([t["doppler_bin"] for t in tg], dtype=np.int32)
        out[f"f{k}_antenna"] = np.array([t["antenna"] for t in tg], dtype=np.int32)
        out[f"f{k}_azimuth_deg"] = np.array([t["azimuth_deg"] for t in tg])
        out[f"f{k}_confidence"] = np.array([t["confidence"] for t in tg])
        out[f"f{k}_power_db"] = np.array([t["power_db"] for t in tg])
        print(f"  robust frame {k}: {len(tg)} reliable targets")
    np.savez_compressed(os.path.join(OUT, "robust_3frames.npz"), **out)


def main():
    ref = ref_import.load()
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    for name, c in CASES.items():
        if only and name not in only:
            continue
        run_case(name, c, ref)
    if not only or "robust" in only:
        run_robust(ref)


if __name__ == "__main__":
    main()
