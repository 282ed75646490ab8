"""Detection fused into the persistent 2-D FFT kernel (rs_range_doppler_detect, csrc/rs_fft2d_ws.cu DETECT +
csrc/rs_detect.cu compact_masks_kernel) against the two-stage path (rs_range_doppler_fft + rs_detect) and the oracle
(dechirp.py:215-278): segments, order, leaders, flags and counters must be identical bit for bit."""
import numpy as np
import pytest
import torch

from oracle import radar_oracle as orc

pytestmark = pytest.mark.gpu

ENV = ("RS_FUSED_FFT", "RS_FUSED_NC", "RS_SPLIT_FFT", "RS_K12", "RS_K12_STORE", "RS_K12_CLUSTERS", "RS_K12_STRICT",
       "RS_K12_VARIANT", "RS_K12_SIDE", "RS_FUSED_DETECT", "RS_SPLIT_DETECT")


def _pipe(A, monkeypatch, env, **kw):
    from radar_slam_b200 import RadarConfig, FramePipeline
    for k in ENV:
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    return FramePipeline(RadarConfig(chirp_duration=25.6e-6, num_chirps=128, num_antennas=A, **kw))


def _noise_cube(F, A, seed, amp=0.1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    re = torch.randn((F, A, 128, 256, 2), generator=g, device="cuda", dtype=torch.float32) * amp
    return torch.view_as_complex(re).contiguous()


def _fields(det):
    n = det.F * det.ntiles
    cnt = det.count[:n].cpu().numpy()
    nl = det.nlead[:n].cpu().numpy()
    out = {"count": cnt, "nlead": nl, "nnear": det.nnear[:n].cpu().numpy(), "overflow": det.overflow[:det.F].cpu().numpy(),
           "psum": det.psum[:n].cpu().numpy()}
    slot = np.arange(det.seg_cap)[None, :]
    m, ml = slot < cnt[:, None], slot < nl[:, None]
    for name in ("key", "power", "flags"):
        out[name] = getattr(det, name)[: n * det.seg_cap].view(n, det.seg_cap).cpu().numpy()[m]
    out["lead"] = det.lead[: n * det.seg_cap].view(n, det.seg_cap).cpu().numpy()[ml]
    return out


def _compare(a, b, what):
    for k in ("count", "nlead", "nnear", "overflow", "key", "flags", "lead", "power"):
        assert np.array_equal(a[k], b[k]), (what, k)
    assert np.allclose(a["psum"], b["psum"], rtol=1e-5), what          # same terms, another summation order


@pytest.mark.parametrize("A,F,kw", [
    (8, 3, {}),                                              # the benchmark shape
    (16, 2, {}),                                             # configs[4]: two antenna octets per range tile
    (8, 2, {"det_eps": 2e-2}),                               # thousands of cells inside the guard band: the exact rule + near / cand masks
    (8, 2, {"min_range": 6.0, "max_range": 30.0}),           # range gate
    (8, 2, {"threshold_db": 30.0}),                          # few detections
    (8, 70, {}),                                             # >= 64 frames: the last frames take the side kernels
])
def test_fused_equals_two_stage_path(monkeypatch, A, F, kw):
    cube = _noise_cube(F, A, 11 + A + F)
    strict = {"RS_K12": "ws", "RS_K12_STRICT": "1"}
    pipe = _pipe(A, monkeypatch, strict, **kw)
    rds_f, det_f = pipe.range_doppler_detect(cube)
    fused = _fields(det_f)
    rds_s = pipe.range_doppler(cube)
    split = _fields(pipe.detect(rds_s))
    torch.cuda.synchronize()
    assert torch.equal(rds_f, rds_s)
    assert fused["key"].size > 100
    if "det_eps" in kw:
        assert fused["nnear"].sum() > 1000 and (fused["flags"] & 16).sum() > 100
    _compare(fused, split, "default")
    # few clusters: every cluster walks many planes, both M buffers and both halo buffers wrap many times
    for env in (dict(strict, RS_K12_CLUSTERS="1"), dict(strict, RS_K12_CLUSTERS="3", RS_K12_SIDE="300"),
                dict(strict, RS_FUSED_DETECT="0")):
        pipe2 = _pipe(A, monkeypatch, env, **kw)
        rds2, det2 = pipe2.range_doppler_detect(cube)
        torch.cuda.synchronize()
        assert torch.equal(rds2, rds_s), env
        _compare(_fields(det2), split, str(env))


def test_fused_detection_matches_oracle(monkeypatch):
    """Targets + noise, detection keys against the oracle's extract_peaks on the oracle's fp64 RDS (every disagreement
    must carry RS_FLAG_NEARMAX), then exact after the fp64 recheck."""
    A = 8
    p = orc.RadarParams(chirp_duration=25.6e-6, num_chirps=128, num_antennas=A)
    scene = np.array([(8.0, 0.0, -10.0, 0.0), (12.0, 0.5, -8.0, 0.0), (25.0, -0.7, 0.0, 0.0)])
    np.random.seed(77)
    cube = np.stack([orc.synthesize_frame(p, scene) for _ in range(2)]).astype(np.complex64)
    pipe = _pipe(A, monkeypatch, {"RS_K12": "ws", "RS_K12_STRICT": "1"})
    dev = torch.from_numpy(cube).cuda()
    rds, det = pipe.range_doppler_detect(dev)
    pipe.recheck_detections(dev, det)
    torch.cuda.synchronize()
    for f in range(2):
        ref = orc.range_doppler_spectrum(cube[f].astype(np.complex128), p)
        pk = orc.extract_peaks(ref, p, threshold_db=-20.0)
        want = (pk["antenna"].astype(np.uint32) << 24) | (pk["range_bin"].astype(np.uint32) << 12) | pk["doppler_bin"].astype(np.uint32)
        got = det.frame(f)
        assert np.array_equal(got["key"], want)
        pw = np.abs(ref[pk["antenna"], pk["range_bin"], pk["doppler_bin"]]) ** 2
        assert np.allclose(got["power"], pw, rtol=1e-4)


def test_process_uses_the_fused_entry(monkeypatch):
    """FramePipeline.process calls rs_range_doppler_detect (one stage call instead of two) and gives the same velocities
    as the two-stage path."""
    A = 8
    cube = _noise_cube(4, A, 5)
    pipe = _pipe(A, monkeypatch, {})
    pipe.call_counts = {}
    v1 = pipe.process(cube).clone()
    assert pipe.call_counts.get("rs_range_doppler_detect") == 1 and "rs_detect" not in pipe.call_counts
    pipe2 = _pipe(A, monkeypatch, {"RS_SPLIT_DETECT": "1"})
    pipe2.call_counts = {}
    v2 = pipe2.process(cube)
    assert pipe2.call_counts.get("rs_detect") == 1
    torch.cuda.synchronize()
    assert torch.equal(v1, v2)


def test_process_without_join_pipelines_across_calls(monkeypatch):
    """process(join=False): the recheck + solve of a call stay on the side stream (workspace sets alternate across calls)
    until the next call or join(); results equal the joined calls'."""
    A = 8
    cubes = [_noise_cube(3, A, 40 + i) for i in range(5)]
    pipe = _pipe(A, monkeypatch, {})
    want = [pipe.process(c).clone() for c in cubes]
    outs = [torch.full((3, 8), float("nan"), dtype=torch.float64, device="cuda") for _ in cubes]
    for c, o in zip(cubes, outs):
        pipe.process(c, vel_out=o, join=False)
    pipe.join()
    torch.cuda.synchronize()
    for w, o in zip(want, outs):
        assert torch.equal(w, o)
    # a joined call after unjoined ones, and the non-overlapped (keep) path, flush what is pending
    pipe.process(cubes[0], vel_out=outs[0], join=False)
    v, _, _ = pipe.process(cubes[1], keep=True)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], want[0]) and torch.equal(v, want[1])


@pytest.mark.parametrize("A,method,env", [(8, "music", {}), (16, "music", {}), (8, "music", {"RS_ANGLES_TC": "1"}),
                                          (8, "esprit", {}), (8, "beamforming", {"RS_ANGLES_MMA": "0"})])
def test_power_written_by_the_angle_stage(monkeypatch, A, method, env):
    """range_doppler_detect(defer_power=True) leaves det.power to rs_angles (det_power_out): the tensor-core scans write
    it from the snapshot registers, the other paths gather it -- bit-identical to rs_detect's values either way."""
    cube = _noise_cube(2, A, 3 + A)
    pipe = _pipe(A, monkeypatch, dict(env, RS_K12="ws", RS_K12_STRICT="1"), method=method, search_resolution=1.0)
    ref = None
    for how in ("angles", "on demand"):
        rds, det = pipe.range_doppler_detect(cube, defer_power=True)
        assert det.power_pending
        det.power.fill_(-7.0)
        pipe.angles(rds, det, write_power=(how == "angles"))
        assert det.power_pending == (how != "angles")
        det.materialize_power()
        assert not det.power_pending
        ref = _fields(pipe.detect(rds)) if ref is None else ref
        torch.cuda.synchronize()
        a = _fields(det)
        assert np.array_equal(a["key"], ref["key"]) and np.array_equal(a["power"], ref["power"]), how
        assert a["power"].min() > 0
