"""the stage hand-over of the reference's pipeline script -- np.savez(peaks) / np.load / process_targets /
np.savez(targets) / np.load / solve_velocity (run_ego_motion_pipeline.py:155-169, 207-219, 259-275) -- with the
column-backed records, against the same classes fed plain lists of dicts."""
import numpy as np
import pytest

from golden_util import load_case, params_of, make_input

pytestmark = pytest.mark.gpu


def _eq(a, b):
    return (a is None and b is None) or np.array_equal(np.asarray(a), np.asarray(b))


@pytest.mark.parametrize("name,method", [("c4_sparse", "music"), ("small_hamming_nodc", "esprit")])
def test_stage_files_round_trip_without_dicts(tmp_path, name, method):
    from src.radar_signal.dechirp import SignalPreprocessor
    from src.angle_estimation.angle_estimation import AngleEstimator
    from src.velocity_solver.velocity_solver import VelocitySolver
    from radar_slam_b200.compat.lazy import LazyRecords, records_of
    g, cfg = load_case(name)
    p = params_of(cfg)
    cube = make_input(cfg).astype(np.complex128)
    pre = SignalPreprocessor(fc=p.fc, bandwidth=p.bandwidth, chirp_duration=p.chirp_duration, pri=p.pri,
                             num_chirps=p.num_chirps, sampling_rate=p.sampling_rate, window_type=p.window_type,
                             dc_removal=p.dc_removal)
    est = AngleEstimator(fc=p.fc, antenna_spacing=p.spacing, num_antennas=p.num_antennas)
    sol = VelocitySolver(fc=p.fc, lambda_c=p.lambda_c, num_antennas=p.num_antennas, antenna_spacing=p.spacing)

    # step 2 of the script
    rds = pre.generate_range_doppler_spectrum(cube)
    info = pre.extract_range_doppler_peaks(rds, threshold_db=cfg["thr"])
    assert isinstance(info["peaks"], LazyRecords) and len(info["peaks"]) > 10
    np.save(tmp_path / "f_rds.npy", rds)
    np.savez(tmp_path / "f_peaks.npz", **info)
    # step 3
    rds2 = np.load(tmp_path / "f_rds.npy")
    info2 = dict(np.load(tmp_path / "f_peaks.npz", allow_pickle=True))
    assert info2["peaks"].ndim == 0 and isinstance(records_of(info2["peaks"]), LazyRecords)
    sel = np.arange(0, len(info["peaks"]), max(1, len(info["peaks"]) // 40))       # keep the bounded solve small
    sub = {**info2, "peaks": records_of(info2["peaks"])[sel]}
    targets = est.process_targets(rds2, sub, method=method)
    assert isinstance(targets, LazyRecords) and len(targets) == len(sel)
    np.savez(tmp_path / "f_angles.npz", targets=targets, radar_params={"fc": p.fc})
    # step 4
    tl = np.load(tmp_path / "f_angles.npz", allow_pickle=True)["targets"]
    res = sol.solve_velocity(rds2, tl, dt=0.1)

    # the same three calls on plain lists of dicts, as the reference's own classes would exchange them
    peaks_list = [info["peaks"][int(i)] for i in sel]
    assert all(isinstance(d, dict) and len(d) == 6 for d in peaks_list)
    targets_list = est.process_targets(rds, {"peaks": peaks_list}, method=method).tolist()
    assert len(targets_list) == len(targets) and set(targets_list[0]) == set(targets[0]) and len(targets_list[0]) == 10
    for a, b in zip(targets, targets_list):
        assert all(_eq(a[k], b[k]) for k in a)
    res_list = sol.solve_velocity(rds, targets_list, dt=0.1)
    assert res["success"] == res_list["success"]
    for k in ("velocity", "angular_velocity", "residuals", "observed_phases"):
        assert np.array_equal(np.asarray(res[k]), np.asarray(res_list[k])), k


def test_cli_file_interfaces_chain(tmp_path):
    """process_frame -> extract_angles_from_rds -> estimate_velocity_from_angles through their FILE interfaces
    (dechirp.py:313-355, angle_estimation.py:368-417, velocity_solver.py:418-467): what the three CLIs do.  The peaks
    written by process_frame are LazyRecords (a 0-d object array once np.load-ed); the angle stage must unwrap them."""
    from src.radar_signal.dechirp import process_frame
    from src.angle_estimation.angle_estimation import extract_angles_from_rds
    from src.robust_angle_estimation import extract_angles_robust
    from src.velocity_solver.velocity_solver import estimate_velocity_from_angles
    from oracle import radar_oracle as orc
    p = orc.RadarParams(chirp_duration=12.8e-6, num_chirps=32)
    np.random.seed(4)
    raw = orc.synthesize_frame(p, np.array([[12.0, 0.4, -3.0, 0.0], [20.0, -0.3, 0.0, 0.0], [30.0, 0.1, -6.0, 0.0]]))
    np.save(tmp_path / "frame.npy", raw)
    params = {"fc": 77e9, "bandwidth": 1e9, "chirp_duration": 12.8e-6, "pri": 100e-6, "num_chirps": 32, "sampling_rate": 10e6}
    out = process_frame(str(tmp_path / "frame.npy"), str(tmp_path / "frame_rds.npy"), params)
    want = orc.extract_peaks(orc.range_doppler_spectrum(raw, p), p, threshold_db=-20.0)
    assert out["num_peaks"] == len(want["antenna"]) > 0
    res = extract_angles_from_rds(str(tmp_path / "frame_rds.npy"), str(tmp_path / "frame_rds_peaks.npz"),
                                  str(tmp_path / "frame_angles.npz"), method="beamforming")
    assert res["num_targets"] == out["num_peaks"]
    rob = extract_angles_robust(str(tmp_path / "frame_rds.npy"), str(tmp_path / "frame_rds_peaks.npz"),
                                str(tmp_path / "frame_robust.npz"))
    assert rob["num_targets"] >= 0 and "statistics" in rob
    vel = estimate_velocity_from_angles(str(tmp_path / "frame_angles.npz"), str(tmp_path / "frame_rds.npy"),
                                        str(tmp_path / "frame_vel.npz"), dt=0.1)
    assert vel["success"] and vel["num_targets"] == res["num_targets"]
    grid = orc.azimuth_grid((-90, 90), 0.5)
    steer = orc.steering_matrix(grid, p.antenna_positions, p.lambda_c)
    sigs = orc.spatial_signatures(orc.range_doppler_spectrum(raw, p), want["range_bin"], want["doppler_bin"])
    _, ang = orc.argmax_angles(orc.beamforming_spectra(sigs, steer), grid)
    ref = orc.solve_velocity(want["range_m"], np.radians(ang), sigs, 3e8 / 77e9, 0.1)
    assert np.abs(vel["velocity"][:2] - ref["velocity"][:2]).max() < 1e-3
