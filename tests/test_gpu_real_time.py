"""the real-time shell (src/core/real_time_processor.py) with the reference's placeholder processing
replaced by the CUDA path: what the worker thread produces equals the direct calls of the drop-in classes."""
import time

import numpy as np
import pytest

from golden_util import load_case, params_of, make_input

pytestmark = pytest.mark.gpu


def test_worker_runs_the_real_angle_and_velocity_stages():
    from src.core.real_time_processor import create_real_time_estimator, RealTimeVelocityEstimator
    from src.radar_signal.dechirp import SignalPreprocessor
    from src.angle_estimation.angle_estimation import AngleEstimator
    from src.velocity_solver.velocity_solver import VelocitySolver
    g, cfg = load_case("c4_sparse")
    p = params_of(cfg)
    cube = make_input(cfg).astype(np.complex128)
    pre = SignalPreprocessor(fc=p.fc, bandwidth=p.bandwidth, chirp_duration=p.chirp_duration, pri=p.pri,
                             num_chirps=p.num_chirps, sampling_rate=p.sampling_rate, window_type=p.window_type,
                             dc_removal=p.dc_removal)
    rds = pre.generate_range_doppler_spectrum(cube)
    info = pre.extract_range_doppler_peaks(rds, threshold_db=cfg["thr"])
    sel = np.arange(0, len(info["peaks"]), max(1, len(info["peaks"]) // 40))
    sub = {**info, "peaks": info["peaks"][sel]}

    radar = {"fc": p.fc, "lambda_c": p.lambda_c, "num_antennas": p.num_antennas, "antenna_spacing": p.spacing}
    est = create_real_time_estimator(radar, frame_buffer_size=4, use_parallel=False)
    assert isinstance(est, RealTimeVelocityEstimator) and est.get_latest_velocity_estimate() is None
    est.start_estimation()
    try:
        ids = [est.add_frame(rds, sub), est.add_frame(rds, sub)]
        assert ids == [0, 1]
        t0 = time.time()
        while len(est.real_time_processor.get_latest_results(2)) < 2 and time.time() - t0 < 120:
            time.sleep(0.05)
        frames = est.real_time_processor.get_latest_results(2)
        assert [f.frame_id for f in frames] == [0, 1]
        got = est.get_latest_velocity_estimate()
    finally:
        est.stop_estimation()

    want_targets = AngleEstimator(fc=p.fc, antenna_spacing=p.spacing, num_antennas=p.num_antennas).process_targets(rds, sub, "music")
    assert len(frames[1].targets) == len(want_targets) == len(sel)
    assert np.array_equal(np.array([t["azimuth_deg"] for t in frames[1].targets]),
                          np.array([t["azimuth_deg"] for t in want_targets]))           # not np.random.uniform(-90, 90)
    want = VelocitySolver(fc=p.fc, lambda_c=p.lambda_c, num_antennas=p.num_antennas,
                          antenna_spacing=p.spacing).solve_velocity(rds, want_targets, dt=0.1)
    assert want["success"] and got["frame_id"] == 1
    assert np.array_equal(got["velocity"], np.asarray(want["velocity"]))
    assert np.array_equal(got["angular_velocity"], np.asarray(want["angular_velocity"]))
    assert 0.0 < got["confidence"] <= 1.0
    m = est.get_estimation_statistics()
    assert m["processing_metrics"]["frames_processed"] == 2 and m["velocity_history_length"] == 1
    assert set(est.real_time_processor.get_system_status()) >= {"cpu_percent", "memory_percent", "processing_metrics"}
