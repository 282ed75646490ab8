"""Device-side frame synthesis (rs_synthesize_frames) against the oracle's restatement of
FMCWRadarSimulator.synthesize_frame (scripts/simulate_raw.py:147-221)."""
import numpy as np
import pytest
import torch

from oracle import radar_oracle as orc

pytestmark = pytest.mark.gpu

SCENE = np.array([(8.0, np.radians(0.0), -10.0, 0.0), (12.0, np.radians(30.0), -8.0, 1.5),
                  (16.0, np.radians(-20.0), -6.0, -3.0), (0.0, 0.1, -3.0, 0.0),          # range 0: skipped (:177)
                  (3000.0, 0.3, 20.0, 0.0),                                              # delay beyond the chirp: no samples
                  (25.0, np.radians(-40.0), 0.0, 0.0), (np.nan, 0.0, 0.0, 0.0)])


def _cfg(p):
    from radar_slam_b200 import RadarConfig
    return RadarConfig(fc=p.fc, bandwidth=p.bandwidth, chirp_duration=p.chirp_duration, pri=p.pri,
                       num_chirps=p.num_chirps, sampling_rate=p.sampling_rate, num_antennas=p.num_antennas)


@pytest.mark.parametrize("S,C,A", [(256, 16, 8), (400, 4, 3), (512, 2, 16), (25, 3, 2)])
def test_scatterer_term_matches_oracle(S, C, A):
    from radar_slam_b200 import synth
    p = orc.RadarParams(chirp_duration=S / 10e6, num_chirps=C, num_antennas=A, noise_power=0.0)
    want = orc.scatterer_response(p, SCENE[np.isfinite(SCENE).all(axis=1)])             # [A, S] complex128
    got = synth.synthesize_frames(_cfg(p), SCENE, 2, seed=1, noise_power=0.0).cpu().numpy()
    assert got.shape == (2, A, C, S)
    scale = np.abs(want).max()
    assert np.abs(got[0] - want[:, None, :]).max() <= 2e-7 * scale                       # complex64 rounding of the plane
    assert np.array_equal(got[0], got[1]) and np.array_equal(got[0][:, 0], got[0][:, -1])   # chirp invariant (:190-209)
    # and the host plane used by the CPU arms is the same thing
    assert np.abs(synth.scatterer_term(_cfg(p), SCENE) - want).max() <= 1e-12 * scale


def test_noise_is_gaussian_keyed_by_frame():
    from radar_slam_b200 import synth
    p = orc.RadarParams(chirp_duration=25.6e-6, num_chirps=64, num_antennas=8)
    cfg = _cfg(p)
    none = np.zeros((0, 4))
    a = synth.synthesize_frames(cfg, none, 4, seed=7, noise_power=0.01)
    b = synth.synthesize_frames(cfg, none, 2, seed=7, noise_power=0.01, first_frame=2)
    c = synth.synthesize_frames(cfg, none, 1, seed=8, noise_power=0.01)
    assert torch.equal(a[2:], b)                          # frame k depends on (seed, k) only: any rank can make its shard
    assert not torch.equal(a[0], a[1]) and not torch.equal(a[0], c[0])
    z = torch.view_as_real(a).double().cpu().numpy().reshape(-1, 2) / 0.1     # unit normals if the model is right
    n = len(z)
    assert abs(z.mean()) < 5 / np.sqrt(2 * n)
    assert abs(z.var() - 1) < 5 * np.sqrt(2 / (2 * n))
    assert abs(np.mean(z[:, 0] * z[:, 1])) < 5 / np.sqrt(n)                   # real and imaginary parts uncorrelated
    assert abs(np.mean(z[:-1, 0] * z[1:, 0])) < 5 / np.sqrt(n)                # neighbouring samples uncorrelated
    assert abs(np.mean(z ** 4) - 3) < 0.05                                     # kurtosis of a normal
    assert 4.5 < np.abs(z).max() < 7                                           # tails present, nothing absurd
    # an odd number of samples per chirp takes the scalar kernel: same stream, same statistics
    p_odd = orc.RadarParams(chirp_duration=25.5e-6, num_chirps=64, num_antennas=8)
    zo = torch.view_as_real(synth.synthesize_frames(_cfg(p_odd), none, 2, seed=7, noise_power=0.01)).double().cpu().numpy().reshape(-1, 2) / 0.1
    assert p_odd.samples_per_chirp % 2 == 1 and abs(zo.mean()) < 5 / np.sqrt(2 * len(zo)) and abs(zo.var() - 1) < 0.02
    # Kolmogorov-Smirnov against the normal CDF on a subsample
    from scipy import stats
    assert stats.kstest(z[::97, 0], "norm").pvalue > 1e-3


def test_simulator_class_matches_reference_interface():
    import pandas as pd
    from radar_slam_b200.compat.simulate_raw import FMCWRadarSimulator
    sim = FMCWRadarSimulator(chirp_duration=12.8e-6, num_chirps=8, num_antennas=4, noise_power=0.0)
    df = pd.DataFrame({"range_sc": [10.0, 14.0], "azimuth_sc": [0.2, -0.3], "rcs": [-5.0, 0.0], "vr": [0.0, 2.0]})
    frame = sim.synthesize_frame(df, frame_idx=3)
    assert frame.shape == (4, 8, 128) and frame.dtype == np.complex128
    p = orc.RadarParams(chirp_duration=12.8e-6, num_chirps=8, num_antennas=4, noise_power=0.0)
    want = orc.scatterer_response(p, df.to_numpy()[:, :4])
    assert np.abs(frame - want[:, None, :]).max() <= 2e-7 * np.abs(want).max()
    td, ph = sim.compute_target_response(10.0, 0.2, 0.0, -5.0, 0.0)
    assert td == 2 * 10.0 / 3e8 and ph.shape == (4,)
    with pytest.raises(FileNotFoundError):
        sim.process_sequence("/nonexistent/sequence", "/tmp/rs_b200_unused_out")


def test_process_sequence_file_interface(tmp_path, monkeypatch):
    """simulate_raw.py:223-335: RadarScenes detections grouped by timestamp -> frame_NNNN.npy + synthesis_metadata.json.
    h5py is replaced by a stand-in that serves a structured array (the container has no HDF5 library)."""
    import json
    import sys
    import types
    from radar_slam_b200.compat.simulate_raw import FMCWRadarSimulator
    rng = np.random.RandomState(4)
    n = 23
    rec = np.zeros(n, dtype=[("timestamp", "i8"), ("range_sc", "f4"), ("azimuth_sc", "f4"), ("rcs", "f4"), ("vr", "f4"),
                             ("x_cc", "f4"), ("y_cc", "f4")])
    rec["timestamp"] = rng.choice([100, 250, 170, 900], size=n)           # unsorted: frames are the SORTED unique stamps
    rec["range_sc"] = rng.uniform(5, 40, n)
    rec["range_sc"][[3, 11]] = 0.0                                         # invalid scatterers are counted and skipped
    rec["azimuth_sc"] = rng.uniform(-1, 1, n)
    rec["rcs"] = rng.uniform(-10, 5, n)
    rec["vr"] = rng.uniform(-3, 3, n)
    seq = tmp_path / "sequence_1"
    seq.mkdir()
    (seq / "scenes.json").write_text("{}")
    np.save(seq / "radar_data.h5.npy", rec)
    (seq / "radar_data.h5").write_bytes(b"stand-in")

    class FakeFile:
        def __init__(self, path, mode="r"):
            self.data = {"radar_data": np.load(str(path) + ".npy")}
        def __enter__(self):
            return self.data
        def __exit__(self, *a):
            return False
    monkeypatch.setitem(sys.modules, "h5py", types.SimpleNamespace(File=FakeFile))
    sim = FMCWRadarSimulator(chirp_duration=12.8e-6, num_chirps=4, num_antennas=4, noise_power=0.0)
    out = tmp_path / "raw"
    stats = sim.process_sequence(str(seq), str(out), max_frames=3, batch_frames=2)
    stamps = np.unique(rec["timestamp"])[:3]
    sel = np.isin(rec["timestamp"], stamps)
    assert stats == {"total_frames": 3, "processed_frames": 3, "total_scatterers": int(sel.sum()),
                     "valid_scatterers": int((rec["range_sc"][sel] > 0).sum())}
    p = orc.RadarParams(chirp_duration=12.8e-6, num_chirps=4, num_antennas=4, noise_power=0.0)
    for k, ts in enumerate(stamps):
        rows = rec[rec["timestamp"] == ts]
        table = np.stack([rows[c].astype(np.float64) for c in ("range_sc", "azimuth_sc", "rcs", "vr")], axis=1)
        want = orc.scatterer_response(p, table)
        got = np.load(out / f"frame_{k:04d}.npy")
        assert got.shape == (4, 4, 128) and got.dtype == np.complex128
        assert np.abs(got - want[:, None, :]).max() <= 2e-7 * max(np.abs(want).max(), 1e-30)
    meta = json.loads((out / "synthesis_metadata.json").read_text())
    assert meta["processing_stats"] == stats and meta["radar_params"]["num_chirps"] == 4
    assert not (out / "frame_0003.npy").exists()
