"""Parity at the shapes of BASELINE.json configs[2..4] and at the edges the reference exercises:
192 virtual channels / 512x256 cube / ESPRIT, the dense-scene threshold, 16 channels, non-power-of-two
and odd sizes, asymmetric search grids, empty and overflowing detection lists."""
import numpy as np
import pytest
import torch

from oracle import radar_oracle as orc

pytestmark = pytest.mark.gpu

SCENE = np.array([(8.0, np.radians(0.0), -10.0, 0.0), (12.0, np.radians(30.0), -8.0, 0.0),
                  (16.0, np.radians(-20.0), -6.0, 0.0), (20.0, np.radians(10.0), -3.0, 0.0),
                  (25.0, np.radians(-40.0), 0.0, 0.0)])


def _pipe(p, **kw):
    from radar_slam_b200 import RadarConfig, FramePipeline
    return FramePipeline(RadarConfig(fc=p.fc, bandwidth=p.bandwidth, chirp_duration=p.chirp_duration, pri=p.pri,
                                     num_chirps=p.num_chirps, sampling_rate=p.sampling_rate,
                                     window_type=p.window_type, dc_removal=p.dc_removal,
                                     num_antennas=p.num_antennas, **kw))


def _keys(pk):
    return (pk["antenna"].astype(np.uint32) << 24) | (pk["range_bin"].astype(np.uint32) << 12) | pk["doppler_bin"].astype(np.uint32)


def _run(p, cube64, thr, method, res=1.0, search_range=(-90, 90)):
    pipe = _pipe(p, threshold_db=thr, method=method, search_resolution=res, search_range=search_range)
    vel, rds, det = pipe.process(torch.from_numpy(cube64[None]).cuda(), keep=True)
    torch.cuda.synchronize()
    ref = orc.range_doppler_spectrum(cube64.astype(np.complex128), p)
    pk = orc.extract_peaks(ref, p, threshold_db=thr)
    return pipe, vel, rds, det, ref, pk


def _check_rds_and_keys(rds, det, ref, pk):
    got = rds[0].permute(1, 0, 2).cpu().numpy()
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
    assert int(det.overflow.sum()) == 0
    d = det.frame(0)
    want = _keys(pk)
    if not np.array_equal(d["key"], want):
        diff = np.setxor1d(d["key"], want)
        assert len(diff) <= 2 and len(want) > 1000, diff          # guard-band cells only (flag checked below)
        hit = np.isin(d["key"], diff)
        assert np.all(d["flags"][hit] & 2)
    return d


def test_config3_mimo_192_channels_esprit():
    """configs[2]: 192 virtual channels (12Tx x 16Rx as a lambda/2 ULA), 512 samples x 256 chirps, ESPRIT."""
    p = orc.RadarParams(chirp_duration=51.2e-6, num_chirps=256, num_antennas=192)
    rs = np.random.RandomState(31)
    cube = (orc.scatterer_response(p, SCENE)[:, None, :] +
            0.1 * (rs.randn(192, 256, 512) + 1j * rs.randn(192, 256, 512))).astype(np.complex64)
    thr = 39.0
    pipe, vel, rds, det, ref, pk = _run(p, cube, thr, "esprit")
    assert 500 < len(pk["antenna"]) < 400000
    d = _check_rds_and_keys(rds, det, ref, pk)
    common, ia, ib = np.intersect1d(d["key"], _keys(pk), return_indices=True)
    sigs = orc.spatial_signatures(ref, pk["range_bin"][ib], pk["doppler_bin"][ib])
    want = orc.esprit_angles(sigs, p.lambda_c, p.spacing)
    assert np.abs(d["adeg"][ia] - want).max() < 0.05
    sol = orc.solve_velocity(pk["range_m"][ib], np.radians(want), sigs, p.lambda_c, 0.1)
    v = vel[0].cpu().numpy()
    assert v[6] == 1.0 and np.abs(v[:2] - sol["velocity"][:2]).max() < 1e-3
    # beamforming on the same detections through the warp-per-detection (A > 16) kernel
    det2 = pipe.detect(rds)
    pipe.angles(rds, det2, method="beamforming")
    torch.cuda.synchronize()
    d2 = det2.frame(0)
    sub = np.arange(0, len(ia), max(1, len(ia) // 400))
    grid = orc.azimuth_grid((-90, 90), 1.0)
    steer = orc.steering_matrix(grid, p.antenna_positions, p.lambda_c)
    spec = orc.beamforming_spectra(sigs[sub], steer)
    idx = np.argmax(spec, axis=1)
    srt = np.sort(spec, axis=1)
    gap = (srt[:, -1] - srt[:, -2]) / srt[:, -1]
    bad = d2["aidx"][ia][sub] != idx
    assert np.all(gap[bad] < 1e-4) and bad.mean() < 0.02


def test_config4_dense_scene_2k_targets_robust_ls():
    """configs[3]: ~2k detections/frame (threshold 30.5 dB on the default noise), MUSIC 0.5 deg,
    plain and Huber-reweighted least squares."""
    p = orc.RadarParams(chirp_duration=25.6e-6, num_chirps=128, num_antennas=8)
    np.random.seed(77)
    cube = orc.synthesize_frame(p, SCENE).astype(np.complex64)
    pipe, vel, rds, det, ref, pk = _run(p, cube, 30.5, "music", res=0.5)
    assert 1000 < len(pk["antenna"]) < 4000
    d = _check_rds_and_keys(rds, det, ref, pk)
    grid = orc.azimuth_grid((-90, 90), 0.5)
    steer = orc.steering_matrix(grid, p.antenna_positions, p.lambda_c)
    sigs = orc.spatial_signatures(ref, pk["range_bin"], pk["doppler_bin"])
    idx, ang = orc.argmax_angles(orc.music_spectra(sigs, steer), grid)
    bad = d["aidx"] != idx
    assert np.all(d["flags"][bad] & 5) and bad.mean() < 5e-3
    sol = orc.solve_velocity(pk["range_m"], np.radians(ang), sigs, p.lambda_c, 0.1)
    assert np.abs(vel[0, :2].cpu().numpy() - sol["velocity"][:2]).max() < 1e-3
    # robust reweighting (Huber IRLS): oracle IRLS on the same data
    from radar_slam_b200 import RadarConfig, FramePipeline
    import dataclasses
    rp = FramePipeline(dataclasses.replace(pipe.cfg, irls_iters=3, huber_delta=1.0))
    det3 = rp.detect(rds)
    rp.angles(rds, det3)
    v3 = rp.velocity(det3)[0].cpu().numpy()
    y = orc.observed_phases(sigs)
    c, s = np.cos(np.radians(ang)), np.sin(np.radians(ang))
    k = 4 * np.pi * 0.1 / p.lambda_c
    v = orc.box_ls_2d(c, s, y, k)[0]
    for _ in range(3):
        res = np.abs(y - k * (v[0] * c + v[1] * s))
        w = np.where(res > 1.0, 1.0 / np.maximum(res, 1e-300), 1.0)
        v = orc.box_ls_2d(c, s, y, k, weights=w)[0]
    assert np.abs(v3[:2] - v).max() < 1e-3


def test_config5_sixteen_channels_two_frames():
    """configs[4] shape (256 x 128 x 16), default threshold: dense noise detections, frames independent."""
    p = orc.RadarParams(chirp_duration=25.6e-6, num_chirps=128, num_antennas=16)
    np.random.seed(5)
    cube = orc.synthesize_frame(p, SCENE).astype(np.complex64)
    pipe, vel, rds, det, ref, pk = _run(p, cube, -20.0, "music", res=1.0)
    d = _check_rds_and_keys(rds, det, ref, pk)
    assert len(d["key"]) > 40000
    sub = np.arange(0, len(pk["antenna"]), 23)
    grid = orc.azimuth_grid((-90, 90), 1.0)
    steer = orc.steering_matrix(grid, p.antenna_positions, p.lambda_c)
    sigs = orc.spatial_signatures(ref, pk["range_bin"][sub], pk["doppler_bin"][sub])
    idx = np.argmax(orc.beamforming_spectra(sigs, steer), axis=1)
    pos = np.searchsorted(d["key"], _keys(pk)[sub])
    bad = d["aidx"][pos] != idx
    assert np.all(d["flags"][pos][bad] & 5) and bad.mean() < 5e-3


@pytest.mark.parametrize("S,C,A,win", [(400, 32, 8, "hann"), (100, 12, 3, "hamming"), (96, 20, 5, "blackman"),
                                       (64, 64, 2, "hann"), (250, 7, 4, "hann"), (32, 32, 16, "hann")])
def test_odd_and_non_power_of_two_sizes(S, C, A, win):
    """the reference's defaults are S=400; chirp_subset yields arbitrary C; any A >= 2."""
    p = orc.RadarParams(chirp_duration=S / 10e6, num_chirps=C, num_antennas=A, window_type=win)
    np.random.seed(S + C + A)
    cube = orc.synthesize_frame(p, SCENE[:3]).astype(np.complex64)
    pipe, vel, rds, det, ref, pk = _run(p, cube, 12.0, "music", res=2.0)
    got = rds[0].permute(1, 0, 2).cpu().numpy()
    assert np.abs(got - ref).max() <= 3e-6 * np.abs(ref).max()
    d = det.frame(0)
    assert np.array_equal(d["key"], _keys(pk))
    grid = orc.azimuth_grid((-90, 90), 2.0)
    steer = orc.steering_matrix(grid, p.antenna_positions, p.lambda_c)
    sigs = orc.spatial_signatures(ref, pk["range_bin"], pk["doppler_bin"])
    idx = np.argmax(orc.beamforming_spectra(sigs, steer), axis=1)
    bad = d["aidx"] != idx
    assert np.all(d["flags"][bad] & 5)
    if len(pk["antenna"]) >= 3:
        sol = orc.solve_velocity(pk["range_m"], np.radians(grid[idx]), sigs, p.lambda_c, 0.1)
        if sol["well_conditioned"] and not bad.any():
            assert np.abs(vel[0, :2].cpu().numpy() - sol["velocity"][:2]).max() < 1e-3


def test_asymmetric_search_grid_and_spacing():
    """A grid that is not symmetric about 0 takes the generic scan; a non-default antenna spacing changes phi."""
    p = orc.RadarParams(chirp_duration=12.8e-6, num_chirps=32, num_antennas=8, antenna_spacing=0.6 * 3e8 / 77e9)
    np.random.seed(12)
    cube = orc.synthesize_frame(p, SCENE).astype(np.complex64)
    from radar_slam_b200 import RadarConfig, FramePipeline
    pipe = FramePipeline(RadarConfig(chirp_duration=12.8e-6, num_chirps=32, num_antennas=8,
                                     antenna_spacing=p.antenna_spacing, search_range=(-60, 45.5),
                                     search_resolution=0.5, threshold_db=15.0, method="beamforming"))
    vel, rds, det = pipe.process(torch.from_numpy(cube[None]).cuda(), keep=True)
    d = det.frame(0)
    ref = orc.range_doppler_spectrum(cube.astype(np.complex128), p)
    pk = orc.extract_peaks(ref, p, threshold_db=15.0)
    assert np.array_equal(d["key"], _keys(pk))
    grid = orc.azimuth_grid((-60, 45.5), 0.5)
    assert not np.array_equal(grid[::-1], -grid)
    steer = orc.steering_matrix(grid, p.antenna_positions, p.lambda_c)
    sigs = orc.spatial_signatures(ref, pk["range_bin"], pk["doppler_bin"])
    idx = np.argmax(orc.beamforming_spectra(sigs, steer), axis=1)
    bad = d["aidx"] != idx
    assert np.all(d["flags"][bad] & 1) and bad.mean() < 5e-3
    assert np.allclose(d["adeg"][~bad], grid[idx[~bad]])


def test_empty_and_sparse_detection_lists():
    """Threshold above every cell -> no detections -> success False (velocity_solver.py:202-204); exactly two -> False."""
    p = orc.RadarParams(chirp_duration=12.8e-6, num_chirps=32, num_antennas=8)
    np.random.seed(3)
    cube = orc.synthesize_frame(p, SCENE).astype(np.complex64)
    pipe, vel, rds, det, ref, pk = _run(p, cube, 200.0, "music")
    assert len(pk["antenna"]) == 0 and len(det.frame(0)["key"]) == 0
    v = vel[0].cpu().numpy()
    assert v[6] == 0.0 and v[7] == 0.0 and np.all(v[:6] == 0.0)
    # a zero cube: every cell is a plateau tie at -120 dB; nothing above -20 dB
    zero = torch.zeros((1, 8, 32, 128), dtype=torch.complex64, device="cuda")
    vel0, rds0, det0 = pipe.process(zero, keep=True)
    assert int(det0.per_frame_counts().sum()) == 0 and float(rds0.abs().max()) == 0.0


def test_plateau_overflow_is_reported_and_recoverable():
    """A constant plane makes every gated cell a local maximum (ties count, dechirp.py:251); the default
    segment capacity overflows, the flag is raised, and a full-capacity pipeline returns the reference's list."""
    from radar_slam_b200 import RadarConfig, FramePipeline, _lib
    p = orc.RadarParams(chirp_duration=6.4e-6, num_chirps=32, num_antennas=8)
    cfg = RadarConfig(chirp_duration=6.4e-6, num_chirps=32, num_antennas=8, threshold_db=-20.0)
    pipe = FramePipeline(cfg)
    rds_ref = np.ones((8, 64, 32), dtype=np.complex64)
    rds = torch.from_numpy(np.ascontiguousarray(rds_ref.transpose(1, 0, 2))[None]).cuda()      # [1, R, A, D]
    det = pipe.detect(rds)
    assert int(det.overflow.sum()) == 1
    tr, td, nt = _lib.detect_tiling(64, 32, 8)
    big = FramePipeline(cfg, seg_cap=tr * td * 8)
    det2 = big.detect(rds)
    assert int(det2.overflow.sum()) == 0
    pk = orc.extract_peaks(rds_ref.astype(np.complex128), p, threshold_db=-20.0)
    assert np.array_equal(det2.frame(0)["key"], _keys(pk)) and len(pk["antenna"]) > 8 * 50 * 32


@pytest.mark.parametrize("S,C,A,res", [(128, 64, 16, 1.0),      # pair mode on 16 x 64 tiles
                                       (256, 256, 16, 1.0),     # pair mode, two Doppler tiles per range row
                                       (120, 32, 16, 2.0),      # range bins not a multiple of the tile: per-segment scan
                                       (256, 128, 16, 0.5),     # 361 grid points: six jobs per tile
                                       (128, 64, 12, 1.0),      # padded to 16 channels, no pair mode
                                       (128, 64, 7, 1.0),       # padded to 8 channels
                                       (128, 64, 8, 5.0),       # 37 grid points: a single job per tile
                                       (128, 64, 16, 0.25)])    # 721 grid points: twelve jobs, one CTA per SM
def test_tcgen05_scan_geometries(S, C, A, res):
    """The persistent tcgen05 scan (and its 16-channel pair mode) on dense noise lists at shapes that move its tile / queue
    / job geometry: every disagreement with the oracle's fp64 argmax is flagged, velocity within the contract."""
    p = orc.RadarParams(chirp_duration=S / 10e6, num_chirps=C, num_antennas=A)
    np.random.seed(S + C + A)
    cube = orc.synthesize_frame(p, SCENE[:3]).astype(np.complex64)
    pipe, vel, rds, det, ref, pk = _run(p, cube, -20.0, "music", res=res)
    d = _check_rds_and_keys(rds, det, ref, pk)
    assert len(d["key"]) > 2000
    grid = orc.azimuth_grid((-90, 90), res)
    steer = orc.steering_matrix(grid, p.antenna_positions, p.lambda_c)
    pos = np.searchsorted(d["key"], _keys(pk))
    ok = (pos < len(d["key"])) & (d["key"][np.minimum(pos, len(d["key"]) - 1)] == _keys(pk))
    sigs = orc.spatial_signatures(ref, pk["range_bin"][ok], pk["doppler_bin"][ok])
    idx = np.argmax(orc.beamforming_spectra(sigs, steer), axis=1)
    bad = d["aidx"][pos[ok]] != idx
    assert np.all(d["flags"][pos[ok]][bad] & 5) and bad.mean() < 5e-3
    sol = orc.solve_velocity(pk["range_m"][ok], np.radians(grid[idx]), sigs, p.lambda_c, 0.1)
    assert np.abs(vel[0, :2].cpu().numpy() - sol["velocity"][:2]).max() < 1e-3
