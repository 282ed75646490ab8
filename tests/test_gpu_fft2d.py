"""The fused cluster 2-D FFT (rs_range_doppler_fft, csrc/rs_fft2d.cu and csrc/rs_fft2d_ws.cu) against the oracle and against the two-kernel path:
every cluster size, a chirp subset that still gives a 128-chirp plane, several frames and antennas, DC removal off."""
import numpy as np
import pytest
import torch

from oracle import radar_oracle as orc

pytestmark = pytest.mark.gpu


def _rds(cube, p, env, monkeypatch, subset=None):
    from radar_slam_b200 import RadarConfig, FramePipeline
    for k in ("RS_FUSED_FFT", "RS_FUSED_NC", "RS_SPLIT_FFT", "RS_K12", "RS_K12_STORE", "RS_K12_CLUSTERS", "RS_K12_STRICT", "RS_K12_VARIANT", "RS_K12_SIDE"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    pipe = FramePipeline(RadarConfig(chirp_duration=p.chirp_duration, num_chirps=cube.shape[2], num_antennas=p.num_antennas,
                                     window_type=p.window_type, dc_removal=p.dc_removal))
    out = pipe.range_doppler(torch.from_numpy(cube).cuda(), chirp_subset=subset)
    torch.cuda.synchronize()
    return out.permute(0, 2, 1, 3).cpu().numpy()                 # [F, S, A, C] -> [F, A, S, C]


@pytest.mark.parametrize("A,F,win,dc", [(8, 3, "hann", True), (16, 2, "blackman", True), (3, 2, "hamming", False)])
def test_cluster_kernel_matches_oracle_and_split_path(monkeypatch, A, F, win, dc):
    p = orc.RadarParams(chirp_duration=25.6e-6, num_chirps=128, num_antennas=A, window_type=win, dc_removal=dc)
    scene = np.array([(8.0, 0.0, -10.0, 0.0), (12.0, 0.5, -8.0, 0.0), (25.0, -0.7, 0.0, 0.0)])
    np.random.seed(21 + A)
    cube = np.stack([orc.synthesize_frame(p, scene) for _ in range(F)]).astype(np.complex64)
    ref = np.stack([orc.range_doppler_spectrum(c.astype(np.complex128), p) for c in cube])
    scale = np.abs(ref).max()
    split = _rds(cube, p, {"RS_FUSED_FFT": "0"}, monkeypatch)
    assert np.abs(split - ref).max() <= 2e-6 * scale
    variants = [{"RS_K12": "v1", "RS_FUSED_NC": nc} for nc in ("2", "4", "8")]
    # the persistent warp-specialised TMA kernel: default, bulk-store rows, and so few clusters that every cluster
    # walks many planes (both M buffers and both ring slots wrap several times)
    ws = {"RS_K12": "ws", "RS_K12_STRICT": "1"}                  # strict: fail instead of falling back to v1
    variants += [ws, dict(ws, RS_K12_STORE="tma"), dict(ws, RS_K12_CLUSTERS="1"),
                 dict(ws, RS_K12_STORE="tma", RS_K12_CLUSTERS="2"), dict(ws, RS_K12_VARIANT="0", RS_K12_CLUSTERS="3"),
                 dict(ws, RS_K12_VARIANT="1"), dict(ws, RS_K12_VARIANT="3", RS_K12_STORE="tma", RS_K12_CLUSTERS="1"),
                 dict(ws, RS_K12_VARIANT="3"), dict(ws, RS_K12_VARIANT="2", RS_K12_CLUSTERS="2"),
                 dict(ws, RS_K12_VARIANT="4", RS_K12_CLUSTERS="1"), dict(ws, RS_K12_VARIANT="4", RS_K12_STORE="tma")]
    for env in variants:
        nc = str(env)
        got = _rds(cube, p, env, monkeypatch)
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() <= 2e-6 * scale, nc
        big = np.abs(ref) > 1e-3 * scale
        assert (np.abs(np.abs(got[big]) - np.abs(ref[big])) / np.abs(ref[big])).max() < 1e-4       # BASELINE tolerance
        assert np.abs(got - split).max() <= 1e-6 * scale                                             # same fp32 algorithm
        if dc:
            assert np.all(got[:, :, 128, :] == 0)                # the mean-removed bin (range bin 0, shifted to S/2)


def test_cluster_kernel_on_a_chirp_subset(monkeypatch):
    """dechirp.py:184-187: chirps [16, 144) of a 160-chirp frame are a 256 x 128 plane again -> the cluster kernel."""
    p = orc.RadarParams(chirp_duration=25.6e-6, num_chirps=160, num_antennas=4)
    np.random.seed(5)
    cube = orc.synthesize_frame(p, np.array([(10.0, 0.2, -5.0, 0.0)]))[None].astype(np.complex64)
    ref = orc.range_doppler_spectrum(cube[0].astype(np.complex128), p, chirp_subset=(16, 144))
    got = _rds(cube, p, {}, monkeypatch, subset=(16, 144))[0]
    split = _rds(cube, p, {"RS_FUSED_FFT": "0"}, monkeypatch, subset=(16, 144))[0]
    assert got.shape == ref.shape == (4, 256, 128)
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max() and np.abs(got - split).max() <= 1e-6 * np.abs(ref).max()


def test_side_kernel_share_of_the_frames(monkeypatch):
    """rs_range_doppler_fft sends the last RS_K12_SIDE permille of a batch (>= 64 frames) to a 2-CTA-cluster kernel on a
    forked stream (the SMs the 4-CTA clusters strand): every frame, on either side of the split, must come out right."""
    p = orc.RadarParams(chirp_duration=25.6e-6, num_chirps=128, num_antennas=2)
    rng = np.random.RandomState(3)
    F = 72
    cube = (rng.randn(F, 2, 128, 256) + 1j * rng.randn(F, 2, 128, 256)).astype(np.complex64)
    ref = np.stack([orc.range_doppler_spectrum(c.astype(np.complex128), p) for c in cube])
    scale = np.abs(ref).max()
    first = None
    for side in ("0", "60", "250", "999"):
        got = _rds(cube, p, {"RS_K12": "ws", "RS_K12_STRICT": "1", "RS_K12_SIDE": side}, monkeypatch)
        assert np.abs(got - ref).max() <= 2e-6 * scale, side
        # both kernels run the same f32x2 operations in the same order: where the split falls changes no bit
        first = got if first is None else first
        assert np.array_equal(got, first), side
