"""Pins the oracle restatement against the LIVE reference classes (only where /root/reference
exists, i.e. in the build container; the GPU box relies on tests/golden/ instead)."""
import numpy as np
import pandas as pd
import pytest

from oracle import radar_oracle as orc
from oracle import ref_import
from conftest import c1_scatterers

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")


def _df(sc):
    return pd.DataFrame([{"range_sc": r, "azimuth_sc": a, "rcs": s, "vr": v} for r, a, s, v in sc])


def _params(S=256, C=32, A=8, **kw):
    return orc.RadarParams(chirp_duration=S / 10e6, num_chirps=C, num_antennas=A, **kw)


def _ref_sim(ref, p):
    return ref.FMCWRadarSimulator(fc=p.fc, bandwidth=p.bandwidth, chirp_duration=p.chirp_duration, pri=p.pri,
                                  num_chirps=p.num_chirps, num_antennas=p.num_antennas,
                                  sampling_rate=p.sampling_rate, noise_power=p.noise_power)


def _ref_pre(ref, p):
    return ref.SignalPreprocessor(fc=p.fc, bandwidth=p.bandwidth, chirp_duration=p.chirp_duration, pri=p.pri,
                                  num_chirps=p.num_chirps, sampling_rate=p.sampling_rate,
                                  window_type=p.window_type, dc_removal=p.dc_removal)


@pytest.mark.parametrize("S,C,A", [(256, 32, 8), (400, 16, 4)])
def test_synthesis_bit_exact(S, C, A):
    ref = ref_import.load()
    p = _params(S, C, A)
    sc = c1_scatterers()
    np.random.seed(7)
    want = _ref_sim(ref, p).synthesize_frame(_df(sc))
    np.random.seed(7)
    got = orc.synthesize_frame(p, sc)
    assert got.shape == want.shape == (A, C, S)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("S,C,A,win,dc", [(256, 32, 8, "hann", True), (400, 16, 4, "hamming", True),
                                         (128, 16, 2, "blackman", False)])
def test_rds_bit_exact_and_peaks(S, C, A, win, dc):
    ref = ref_import.load()
    p = _params(S, C, A, window_type=win, dc_removal=dc)
    np.random.seed(11)
    frame = orc.synthesize_frame(p, c1_scatterers())
    pre = _ref_pre(ref, p)
    want = pre.generate_range_doppler_spectrum(frame)
    got = orc.range_doppler_spectrum(frame, p)
    assert got.shape == (A, S, C)
    assert np.array_equal(got, want)
    sub = (3, 11)
    assert np.array_equal(orc.range_doppler_spectrum(frame, p, sub), pre.generate_range_doppler_spectrum(frame, sub))
    for thr in (-20.0, 24.0):
        pi = pre.extract_range_doppler_peaks(want, threshold_db=thr)
        pk = orc.extract_peaks(got, p, threshold_db=thr)
        assert len(pi["peaks"]) == len(pk["antenna"])
        assert [q["antenna"] for q in pi["peaks"]] == pk["antenna"].tolist()
        assert [q["range_bin"] for q in pi["peaks"]] == pk["range_bin"].tolist()
        assert [q["doppler_bin"] for q in pi["peaks"]] == pk["doppler_bin"].tolist()
        assert np.array_equal(np.array([q["power_db"] for q in pi["peaks"]]), pk["power_db"])
        assert np.array_equal(np.array([q["range_m"] for q in pi["peaks"]]), pk["range_m"])
        assert np.array_equal(np.array([q["doppler_hz"] for q in pi["peaks"]]), pk["doppler_hz"])
        assert np.array_equal(pi["power_spectrum_db"], pk["power_spectrum_db"])


def test_unknown_window_raises():
    with pytest.raises(ValueError):
        orc.window("kaiser", 8)


@pytest.mark.parametrize("A,res", [(8, 0.5), (4, 1.0), (16, 2.0)])
def test_angles_match_reference(A, res):
    ref = ref_import.load()
    p = _params(128, 16, A)
    np.random.seed(3)
    frame = orc.synthesize_frame(p, c1_scatterers())
    rds = orc.range_doppler_spectrum(frame, p)
    pk = orc.extract_peaks(rds, p, threshold_db=20.0)
    n = min(40, len(pk["antenna"]))
    est = ref.AngleEstimator(fc=p.fc, num_antennas=A, search_resolution=res)
    grid = orc.azimuth_grid((-90, 90), res)
    assert np.array_equal(grid, est.azimuth_grid)
    steer = orc.steering_matrix(grid, p.antenna_positions, p.lambda_c)
    sigs = orc.spatial_signatures(rds, pk["range_bin"][:n], pk["doppler_bin"][:n])
    spec_m = orc.music_spectra(sigs, steer)
    spec_b = orc.beamforming_spectra(sigs, steer)
    esp = orc.esprit_angles(sigs, p.lambda_c, p.spacing)
    for i in range(n):
        s_ref = est.extract_spatial_signature(rds, pk["range_bin"][i], pk["doppler_bin"][i])
        assert np.array_equal(s_ref, sigs[i])
        a_ref, sp_ref = est.estimate_angle_music(s_ref)
        assert a_ref == grid[np.argmax(spec_m[i])]
        np.testing.assert_allclose(spec_m[i], sp_ref, rtol=1e-9)
        lit = orc.music_spectrum_literal(s_ref, grid, p.antenna_positions, p.lambda_c)
        np.testing.assert_allclose(lit, sp_ref, rtol=1e-12)
        a_ref, sp_ref = est.estimate_angle_beamforming(s_ref)
        assert a_ref == grid[np.argmax(spec_b[i])]
        np.testing.assert_allclose(spec_b[i], sp_ref, rtol=1e-12, atol=1e-15)
        e_ref = est.estimate_angle_esprit(s_ref)
        assert abs(e_ref - esp[i]) < 1e-9
        assert abs(e_ref - orc.esprit_angle_literal(s_ref, p.lambda_c, p.spacing)) < 1e-12


def test_music_guard_zeroes_on_grid_peak():
    """a noise-free on-grid signature makes the denominator ~1e-15, the 1e-12 guard
    zeroes the true peak and argmax returns the lower neighbour."""
    ref = ref_import.load()
    p = _params(128, 16, 8)
    est = ref.AngleEstimator(fc=p.fc, num_antennas=8, search_resolution=1.0)
    grid = orc.azimuth_grid((-90, 90), 1.0)
    steer = orc.steering_matrix(grid, p.antenna_positions, p.lambda_c)
    for deg in (0.0, 30.0, -45.0):
        s = steer[np.argmin(np.abs(grid - deg))] / np.sqrt(8)
        a_ref, _ = est.estimate_angle_music(s)
        idx, ang = orc.argmax_angles(orc.music_spectra(s[None, :], steer), grid)
        assert ang[0] == a_ref


def test_robust_matches_reference_over_frames():
    ref = ref_import.load()
    p = _params(128, 16, 8)
    r_ref = ref.RobustAngleEstimator(fc=p.fc, num_antennas=8, max_targets=25)
    r_orc = orc.RobustOracle(p, max_targets=25)
    pre = _ref_pre(ref, p)
    for k in range(3):
        np.random.seed(100 + k)
        frame = orc.synthesize_frame(p, c1_scatterers())
        rds = pre.generate_range_doppler_spectrum(frame)
        pi = pre.extract_range_doppler_peaks(rds, threshold_db=15.0)
        pk = orc.extract_peaks(rds, p, threshold_db=15.0)
        want = r_ref.process_targets_robust(rds, pi, frame_timestamp=float(k))
        got = r_orc.process(rds, pk, frame_timestamp=float(k))
        assert len(want) == len(got)
        for w, g in zip(want, got):
            assert (w["range_bin"], w["doppler_bin"], w["antenna"]) == (g["range_bin"], g["doppler_bin"], g["antenna"])
            assert abs(w["azimuth_deg"] - g["azimuth_deg"]) < 1e-9
            assert abs(w["confidence"] - g["confidence"]) < 1e-12
            assert w["interference_analysis"]["num_sources"] == g["interference_analysis"]["num_sources"]


@pytest.mark.slow
def test_velocity_ls_matches_differential_evolution():
    """With the correct wavelength the reference's DE converges onto the box-constrained LS point
    in (v_x, v_y); v_z / omega are unobservable and not compared."""
    ref = ref_import.load()
    rng = np.random.RandomState(5)
    N = 12
    az = rng.uniform(-1.2, 1.2, N)
    rm = rng.uniform(5, 40, N)
    lam = 3e8 / 77e9
    k = 4 * np.pi * 0.1 / lam
    vtrue = np.array([0.004, -0.002])
    y = k * (vtrue[0] * np.cos(az) + vtrue[1] * np.sin(az)) + 0.05 * rng.randn(N)
    sigs = np.zeros((N, 8), dtype=complex)
    sigs[:, 0] = 1 / np.sqrt(2)
    sigs[:, 1] = np.exp(1j * y) / np.sqrt(2)
    targets = [{"range_m": rm[i], "azimuth_rad": az[i], "spatial_signature": sigs[i]} for i in range(N)]
    solver = ref.VelocitySolver(lambda_c=lam)
    want = solver.solve_velocity(None, targets, dt=0.1)
    got = orc.solve_velocity(rm, az, sigs, lam, 0.1)
    assert want["success"] and got["success"]
    np.testing.assert_allclose(got["velocity"][:2], want["velocity"][:2], atol=1e-6)
    assert orc.solve_velocity(rm[:2], az[:2], sigs[:2], lam)["success"] is False
    # clipped case: the pipeline script's inverted wavelength (run_ego_motion_pipeline.py:246)
    lam_bad = 77e9 / 3e8
    got_bad = orc.solve_velocity(rm, az, sigs, lam_bad, 0.1)
    assert np.all(np.abs(got_bad["velocity"][:2]) <= 50.0)
