"""The tcgen05 / TMEM angle scan (csrc/rs_angles.cu: angles_tc5_kernel, RS_ANGLES_TC) against the oracle and against the
mma.sync scan: same flag contract (every disagreement with the oracle's argmax carries RS_FLAG_TIE / RS_FLAG_GUARD),
identical results after the fp64 recheck."""
import numpy as np
import pytest
import torch

from oracle import radar_oracle as orc

pytestmark = pytest.mark.gpu


def _run(cube, A, method, tc, recheck, monkeypatch, res=1.0, thr=-20.0):
    from radar_slam_b200 import RadarConfig, FramePipeline
    monkeypatch.setenv("RS_ANGLES_TC", tc)
    cfg = RadarConfig(chirp_duration=25.6e-6, num_chirps=128, num_antennas=A, search_resolution=res, method=method,
                      threshold_db=thr, recheck=recheck)
    pipe = FramePipeline(cfg)
    vel, rds, det = pipe.process(torch.from_numpy(cube).cuda(), keep=True)
    torch.cuda.synchronize()
    return pipe, vel.cpu().numpy(), [det.frame(f) for f in range(cube.shape[0])]


@pytest.mark.parametrize("A,method,res", [(8, "music", 1.0), (16, "music", 1.0), (8, "beamforming", 0.5), (6, "music", 2.0)])
def test_tcgen05_scan_matches_oracle_and_mma(monkeypatch, A, method, res):
    p = orc.RadarParams(chirp_duration=25.6e-6, num_chirps=128, num_antennas=A)
    scene = np.array([(8.0, 0.0, -10.0, 0.0), (12.0, 0.5, -8.0, 0.0), (25.0, -0.7, 0.0, 0.0), (17.0, 1.2, -4.0, 0.0)])
    np.random.seed(40 + A)
    cube = np.stack([orc.synthesize_frame(p, scene) for _ in range(2)]).astype(np.complex64)
    grid = orc.azimuth_grid((-90, 90), res)
    steer = orc.steering_matrix(grid, p.antenna_positions, p.lambda_c)
    # fp32 kernels only: flag contract
    _, _, tc = _run(cube, A, method, "1", False, monkeypatch, res)
    _, _, mma = _run(cube, A, method, "0", False, monkeypatch, res)
    for f in range(2):
        ref = orc.range_doppler_spectrum(cube[f].astype(np.complex128), p)
        assert np.array_equal(tc[f]["key"], mma[f]["key"])
        rb, db = tc[f]["range_bin"], tc[f]["doppler_bin"]
        sigs = orc.spatial_signatures(ref, rb, db)
        idx, _ = orc.argmax_angles(orc.beamforming_spectra(sigs, steer) if method == "beamforming"
                                   else orc.music_spectra(sigs, steer), grid)
        for got in (tc[f], mma[f]):
            bad = got["aidx"] != idx
            assert np.all(got["flags"][bad] & 5), "unflagged angle mismatch"
            assert bad.mean() < 0.02
        # the two scans evaluate the same products (fp16 hi / lo split, fp32 accumulation): they agree outside flagged cells
        diff = tc[f]["aidx"] != mma[f]["aidx"]
        assert np.all((tc[f]["flags"][diff] | mma[f]["flags"][diff]) & 5)
        assert np.allclose(tc[f]["phase"], mma[f]["phase"])
    # with the fp64 recheck both are exact, hence identical, and so are the velocities
    _, v_tc, tc = _run(cube, A, method, "1", True, monkeypatch, res)
    _, v_mma, mma = _run(cube, A, method, "0", True, monkeypatch, res)
    for f in range(2):
        assert np.array_equal(tc[f]["aidx"], mma[f]["aidx"])
    assert np.abs(v_tc[:, :2] - v_mma[:, :2]).max() < 1e-9


def test_pair_mode_equals_per_segment_scan(monkeypatch):
    """16 channels: scanning the two antenna-octet segments of a tile as a pair (every distinct cell once, RS_ANGLES_DEDUP
    default) writes exactly what the per-segment scan writes -- the snapshot of a cell is the same on both octets."""
    from radar_slam_b200 import RadarConfig, FramePipeline
    A = 16
    p = orc.RadarParams(chirp_duration=25.6e-6, num_chirps=128, num_antennas=A)
    scene = np.array([(8.0, 0.0, -10.0, 0.0), (12.0, 0.5, -8.0, 0.0), (25.0, -0.7, 0.0, 0.0)])
    np.random.seed(77)
    cube = torch.from_numpy(np.stack([orc.synthesize_frame(p, scene) for _ in range(3)]).astype(np.complex64)).cuda()
    out = {}
    for dd in ("1", "0"):
        monkeypatch.setenv("RS_ANGLES_TC", "1")
        monkeypatch.setenv("RS_ANGLES_DEDUP", dd)
        cfg = RadarConfig(chirp_duration=25.6e-6, num_chirps=128, num_antennas=A, recheck=False)
        pipe = FramePipeline(cfg)
        rds = pipe.range_doppler(cube)
        det = pipe.angles(rds, pipe.detect(rds))
        torch.cuda.synchronize()
        n = det.F * det.ntiles * det.seg_cap
        m = det.valid_mask().reshape(-1)
        out[dd] = dict(aidx=det.aidx[:n][m].cpu().numpy(), flags=det.flags[:n][m].cpu().numpy(),
                       adeg=det.adeg[:n][m].cpu().numpy(), phase=det.phase[:n][m].cpu().numpy(),
                       ntie=det.ntie.cpu().numpy().copy(), part=det.ls_partials.cpu().numpy().copy(),
                       tie=[np.sort(t[:c]) for t, c in zip(det.tielist.cpu().numpy().reshape(-1, 32),
                                                           np.minimum(det.ntie.cpu().numpy(), 32))])
    a, b = out["1"], out["0"]
    assert a["aidx"].size > 50000
    for k in ("aidx", "flags", "adeg", "phase", "ntie"):
        assert np.array_equal(a[k], b[k]), k
    assert all(np.array_equal(x, y) for x, y in zip(a["tie"], b["tie"]))
    # the seven velocity sums of every segment (slot 8 is padding): same terms, another summation order
    assert np.allclose(a["part"][:, :7], b["part"][:, :7], rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("A,method,res", [(192, "music", 1.0), (40, "beamforming", 0.5), (17, "music", 2.0)])
def test_steering_gemm_on_tensor_cores(monkeypatch, A, method, res):
    """A > 16: the grid scan as a dense [cells x 2A] . [2A x 2G] contraction on tcgen05 (music_tc_kernel) against the
    oracle's fp64 argmax (every disagreement flagged) and against the CUDA-core scan it replaces (RS_MUSIC_TC=0)."""
    from radar_slam_b200 import RadarConfig, FramePipeline
    S, C = 64, 128
    p = orc.RadarParams(chirp_duration=S / 10e6, num_chirps=C, num_antennas=A)
    scene = np.array([(3.0, 0.3, 10.0, 0.0), (6.0, -0.6, 0.0, 0.0)])
    np.random.seed(900 + A)
    cube = np.stack([orc.synthesize_frame(p, scene) for _ in range(2)]).astype(np.complex64)
    grid = orc.azimuth_grid((-90, 90), res)
    steer = orc.steering_matrix(grid, p.antenna_positions, p.lambda_c)
    runs = {}
    for tc in ("1", "0"):
        monkeypatch.setenv("RS_MUSIC_TC", tc)
        cfg = RadarConfig(chirp_duration=S / 10e6, num_chirps=C, num_antennas=A, search_resolution=res, method=method,
                          recheck=False)
        pipe = FramePipeline(cfg)
        vel, rds, det = pipe.process(torch.from_numpy(cube).cuda(), keep=True)
        torch.cuda.synchronize()
        runs[tc] = [det.frame(f) for f in range(2)]
    for f in range(2):
        ref = orc.range_doppler_spectrum(cube[f].astype(np.complex128), p)
        got, old = runs["1"][f], runs["0"][f]
        assert np.array_equal(got["key"], old["key"]) and len(got["key"]) > 5000
        sub = np.arange(0, len(got["key"]), max(1, len(got["key"]) // 4000))
        sigs = orc.spatial_signatures(ref, got["range_bin"][sub], got["doppler_bin"][sub])
        idx, _ = orc.argmax_angles(orc.beamforming_spectra(sigs, steer) if method == "beamforming"
                                   else orc.music_spectra(sigs, steer), grid)
        for r in (got, old):
            bad = r["aidx"][sub] != idx
            assert np.all(r["flags"][sub][bad] & 5), "unflagged angle mismatch"
            assert bad.mean() < 0.03
        diff = got["aidx"] != old["aidx"]
        assert np.all((got["flags"][diff] | old["flags"][diff]) & 5)
        assert np.allclose(got["phase"], old["phase"])
        assert np.mean((got["flags"] & 1) != 0) < 0.05            # the wider TIE band still flags few cells
