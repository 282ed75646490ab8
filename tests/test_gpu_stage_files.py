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
