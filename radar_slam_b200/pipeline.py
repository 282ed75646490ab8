"""Batched device API of the B200 hot path: cube[F,A,C,S] -> RDS -> detections -> angles -> ego-velocity.

This is the layer the drop-in classes under ``src/`` and ``bench.py`` call.  torch owns memory and
streams; all arithmetic runs in libradarslam_b200.so (hand-written sm_100a CUDA).  There is no CPU
fallback: constructing a pipeline without a CUDA device or without the library raises.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib, tables


@dataclass
class RadarConfig:
    """Union of the constructor arguments of the reference's hot-path classes."""
    fc: float = 77e9
    bandwidth: float = 1e9
    chirp_duration: float = 40e-6
    pri: float = 100e-6
    num_chirps: int = 64
    sampling_rate: float = 10e6
    window_type: str = "hann"
    dc_removal: bool = True
    num_antennas: int = 8
    antenna_spacing: Optional[float] = None
    search_range: Tuple[float, float] = (-90, 90)
    search_resolution: float = 0.5
    method: str = "music"
    threshold_db: float = -20.0
    min_range: float = 1.0
    max_range: float = 200.0
    lambda_c_solver: Optional[float] = None      # VelocitySolver(lambda_c=...)
    dt: float = 0.1
    velocity_bound: float = 50.0                 # velocity_solver.py:216
    irls_iters: int = 0                          # robust reweighting (off: the reference has none)
    huber_delta: float = 1.0
    det_eps: float = 2e-6                        # guard bands for decisions fp32 cannot settle
    tie_eps: float = 4e-6
    recheck: bool = True                         # settle the flagged decisions in fp64 from the raw cube
    fft_eps: float = 4e-7                        # 2-norm bound of the fp32 FFT error per element, in units of rms(|X|):
                                                 # measured rms error 1.4e-7 (profiles/fft_error_probe.py), x2.9 margin

    @property
    def lambda_c(self) -> float:
        return tables.C0 / self.fc

    @property
    def samples_per_chirp(self) -> int:
        return int(self.chirp_duration * self.sampling_rate)

    @property
    def chirp_rate(self) -> float:
        return self.bandwidth / self.chirp_duration

    @property
    def range_resolution(self) -> float:
        return tables.C0 / (2 * self.bandwidth)

    @property
    def spacing(self) -> float:
        return self.antenna_spacing or (self.lambda_c / 2)


@dataclass
class Detections:
    """Per-(frame, tile) segments written by rs_detect / rs_angles (layout in radar_slam_b200.h)."""
    key: torch.Tensor
    power: torch.Tensor
    flags: torch.Tensor
    aidx: torch.Tensor
    adeg: torch.Tensor
    phase: torch.Tensor
    count: torch.Tensor
    overflow: torch.Tensor
    seg_cap: int
    ntiles: int
    F: int
    R: int
    D: int
    A: int
    ls_partials: Optional[torch.Tensor] = None     # per-segment normal-equation sums written by rs_angles
    lead: Optional[torch.Tensor] = None            # one leader per distinct cell: position | multiplicity << 16
    nlead: Optional[torch.Tensor] = None
    nnear: Optional[torch.Tensor] = None           # per segment: entries flagged NEARMAX
    psum: Optional[torch.Tensor] = None            # per segment: sum of |X|^2 (noise level for the recheck bound)
    ntie: Optional[torch.Tensor] = None            # per segment: cells flagged TIE / GUARD
    tielist: Optional[torch.Tensor] = None         # per segment: leader indices of the first RS_TIE_LIST_CAP of them
    threshold_db: float = -20.0                    # the threshold rs_detect ran with (needed by the fp64 recheck)
    method: Optional[str] = None                   # the method rs_angles ran with
    tag: str = ""                                  # workspace set the buffers came from
    power_pending: bool = False                    # range_doppler_detect(defer_power=True): `power` not written yet
    power_fill: Optional[object] = None            # callable that writes `power` (gather from the RDS the lists came from)

    def materialize_power(self) -> torch.Tensor:
        """|X|^2 of every entry.  The fused FFT + detection kernel keeps no powers; nothing on the path to the velocity
        reads them, so they are gathered from the RDS when somebody asks (or written by angles(write_power=True))."""
        if self.power_pending:
            self.power_fill()
            self.power_pending = False
        return self.power

    def valid_mask(self) -> torch.Tensor:
        n = self.F * self.ntiles
        slot = torch.arange(self.seg_cap, device=self.key.device, dtype=torch.int32)
        m = slot[None, :] < self.count[:n, None]
        m &= (self.flags[: n * self.seg_cap].view(n, self.seg_cap) & _lib.RS_FLAG_DROPPED) == 0
        return m

    def per_frame_counts(self) -> torch.Tensor:
        return self.valid_mask().view(self.F, -1).sum(dim=1)

    def frame(self, f: int) -> Dict[str, np.ndarray]:
        """Host copy of frame f's detections in the reference's order (antenna, range, doppler)."""
        lo, hi = f * self.ntiles, (f + 1) * self.ntiles
        m = self.valid_mask()[lo:hi].reshape(-1)
        sl = slice(lo * self.seg_cap, hi * self.seg_cap)
        key = self.key[sl][m].to(torch.int64) & 0xFFFFFFFF
        order = torch.argsort(key)
        key = key[order]
        out = {"key": key.cpu().numpy().astype(np.uint32)}
        self.materialize_power()
        for name in ("power", "flags", "aidx", "adeg", "phase"):
            out[name] = getattr(self, name)[sl][m][order].cpu().numpy()
        k = out["key"]
        out["antenna"] = (k >> 24).astype(np.int64)
        out["range_bin"] = ((k >> 12) & 0xFFF).astype(np.int64)
        out["doppler_bin"] = (k & 0xFFF).astype(np.int64)
        return out


class FramePipeline:
    def __init__(self, cfg: RadarConfig, device: Optional[str] = None, seg_cap: Optional[int] = None):
        if not torch.cuda.is_available():
            raise _lib.RadarSlamError("radar_slam_b200 needs a CUDA device (no CPU fallback)")
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device(device or f"cuda:{torch.cuda.current_device()}")
        if cfg.method not in _lib.METHODS:
            raise ValueError(f"Unknown method: {cfg.method}")
        self.seg_cap_override = seg_cap
        self._tab: Dict = {}
        self._ws: Dict = {}
        self.launches = 0          # C-ABI stage calls since the last reset
        self.call_counts = None    # set to {} to count calls per entry point (bench.py: kernels launched)
        self.profile = None        # set to a list to collect (stage, start_event, end_event) per launch
        self.nvtx = os.environ.get("RS_NVTX") == "1"
        self._chunk_no = 0         # chunks processed so far: picks the workspace set
        self._pending = None       # chunk whose fp64 recheck + solve has not been enqueued on the side stream yet

    def _call(self, name: str, *args) -> None:
        fn = getattr(self.lib, name)
        if self.call_counts is not None:
            self.call_counts[name] = self.call_counts.get(name, 0) + 1
        if self.nvtx:                      # RS_NVTX=1: one NVTX range per stage call (nsys / ncu --nvtx timelines)
            torch.cuda.nvtx.range_push(name)
            try:
                _lib.check(fn(*args), name)
            finally:
                torch.cuda.nvtx.range_pop()
            self.launches += 1
            return
        if self.profile is None:
            _lib.check(fn(*args), name)
        else:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(fn(*args), name)
            e1.record()
            self.profile.append((name, e0, e1))
        self.launches += 1

    # ------------------------------------------------------------------ host tables (cached)
    def _dev(self, arr: np.ndarray) -> torch.Tensor:
        return torch.from_numpy(np.ascontiguousarray(arr)).to(self.device)

    def _fft_tables(self, S: int, C: int):
        key = ("fft", S, C)
        if key not in self._tab:
            c = self.cfg
            tab = tables.dechirp_table(c.fc, c.chirp_rate, c.chirp_duration, S, c.window_type)
            self._tab[key] = (self._dev(tab.astype(np.complex64)), self._dev(tables.twiddles(S)),
                              self._dev(tables.twiddles(C)), self._dev(tab),
                              self._dev(tables.twiddles128(S)), self._dev(tables.twiddles128(C)))
        return self._tab[key]

    def _angle_tables(self, A: int):
        key = ("ang", A)
        if key not in self._tab:
            c = self.cfg
            grid = tables.azimuth_grid(c.search_range, c.search_resolution)
            scan, stride = tables.scan_table(grid, c.spacing, c.lambda_c, A)
            pos = np.arange(A) * c.spacing
            steer = tables.steering(grid, pos, c.lambda_c)
            symmetric = bool(np.array_equal(grid[::-1], -grid))
            mma, mma_tiles = (tables.scan_mma_table(scan, len(grid), A) if symmetric and 4 < A <= 16 else (None, 0))
            tc, tc_halves = (tables.scan_tc_table(scan, len(grid), A) if symmetric and 4 < A <= 16 else (None, 0))
            if A > 16:                      # steering GEMM on the tensor cores (MUSIC / beamforming scan at many channels)
                tc, tc_halves = tables.steer_tc_table(steer)
            self._tab[key] = {
                "grid": grid, "G": len(grid), "stride": stride,
                "scan": self._dev(scan), "grid_f32": self._dev(grid.astype(np.float32)),
                "steer64": self._dev(steer.astype(np.complex64)) if A > 16 else None,
                "steer128": self._dev(steer), "grid_cs": self._dev(tables.grid_cos_sin(grid)),
                "symmetric": symmetric, "mma": self._dev(mma) if mma is not None else None, "mma_tiles": mma_tiles,
                "tc": self._dev(tc) if tc is not None else None, "tc_halves": tc_halves,
            }
        return self._tab[key]

    def _gate(self, R: int, min_range: float, max_range: float) -> torch.Tensor:
        key = ("gate", R, float(min_range), float(max_range))
        if key not in self._tab:
            self._tab[key] = self._dev(tables.range_gate(self.cfg.range_resolution, R, min_range, max_range))
        return self._tab[key]

    def _buf(self, name: str, shape, dtype) -> torch.Tensor:
        n = int(np.prod(shape))
        cur = self._ws.get(name)
        if cur is None or cur.dtype != dtype or cur.numel() < n:
            cur = torch.empty(n, dtype=dtype, device=self.device)
            self._ws[name] = cur
        return cur[:n].view(*shape)

    @property
    def stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    # ------------------------------------------------------------------ stages
    def range_doppler(self, cube: torch.Tensor, chirp_subset: Optional[Tuple[int, int]] = None,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """cube complex64 [F, A, C, S] on the device -> RDS complex64 [F, S, A, C'] (range-major planes)."""
        assert cube.is_cuda and cube.dtype == torch.complex64 and cube.dim() == 4 and cube.is_contiguous()
        F, A, C, S = cube.shape
        c0, c1 = (0, C) if chirp_subset is None else chirp_subset
        c0, c1 = max(0, min(C, int(c0))), max(0, min(C, int(c1)))
        Cu = c1 - c0
        if Cu <= 0:
            raise ValueError("empty chirp subset")
        tab, tw_s, tw_c = self._fft_tables(S, Cu)[:3]
        mid = self._buf("mid", (F, S, A, Cu), torch.complex64)
        rds = out if out is not None else torch.empty((F, S, A, Cu), dtype=torch.complex64, device=self.device)
        st = self.stream
        if os.environ.get("RS_SPLIT_FFT") == "1":       # the two stages as separate calls (per-stage timing, tests)
            self._call("rs_range_fft", cube.data_ptr(), tab.data_ptr(), tw_s.data_ptr(), mid.data_ptr(),
                       F, A, C, c0, Cu, S, int(self.cfg.dc_removal), st)
            self._call("rs_doppler_fft", mid.data_ptr(), tw_c.data_ptr(), rds.data_ptr(), F, A, Cu, S, st)
        else:
            self._call("rs_range_doppler_fft", cube.data_ptr(), tab.data_ptr(), tw_s.data_ptr(), tw_c.data_ptr(),
                       mid.data_ptr(), rds.data_ptr(), F, A, C, c0, Cu, S, int(self.cfg.dc_removal), st)
        return rds

    def seg_cap_for(self, R: int, D: int, A: int) -> Tuple[int, int]:
        tr, td, ntiles = _lib.detect_tiling(R, D, A)
        cap = self.seg_cap_override or max(32, (tr * td * min(A, 8) + 3) // 4)
        return cap, ntiles

    def _alloc_detections(self, F: int, R: int, D: int, A: int, threshold_db: Optional[float], workspace) -> Detections:
        c = self.cfg
        cap, ntiles = self.seg_cap_for(R, D, A)
        n = F * ntiles * cap
        tag = workspace if isinstance(workspace, str) else ""      # "0" / "1": double-buffered workspaces
        if workspace:
            self._ws["_nframes" + tag] = F                          # status(): how much of the overflow buffer is current
        alloc = (lambda nm, shp, dt: self._buf(nm + tag, shp, dt)) if workspace else \
            (lambda nm, shp, dt: torch.empty(shp, dtype=dt, device=self.device))
        return Detections(
            key=alloc("det_key", (n,), torch.int32), power=alloc("det_power", (n,), torch.float32),
            flags=alloc("det_flags", (n,), torch.uint8), aidx=alloc("det_aidx", (n,), torch.int32),
            adeg=alloc("det_adeg", (n,), torch.float32), phase=alloc("det_phase", (n,), torch.float32),
            count=alloc("det_count", (F * ntiles,), torch.int32), overflow=alloc("det_overflow", (F,), torch.int32),
            seg_cap=cap, ntiles=ntiles, F=F, R=R, D=D, A=A,
            threshold_db=float(c.threshold_db if threshold_db is None else threshold_db), tag=tag,
            lead=alloc("det_lead", (n,), torch.int32), nlead=alloc("det_nlead", (F * ntiles,), torch.int32),
            nnear=alloc("det_nnear", (F * ntiles,), torch.int32), psum=alloc("det_psum", (F * ntiles,), torch.float32),
            ntie=alloc("det_ntie", (F * ntiles,), torch.int32),
            tielist=alloc("det_tielist", (F * ntiles * _lib.RS_TIE_LIST_CAP,), torch.int32))

    def detect(self, rds: torch.Tensor, threshold_db: Optional[float] = None, min_range: Optional[float] = None,
               max_range: Optional[float] = None, workspace=False) -> Detections:
        assert rds.is_cuda and rds.dtype == torch.complex64 and rds.dim() == 4 and rds.is_contiguous()
        F, R, A, D = rds.shape
        c = self.cfg
        thr = tables.power_threshold(c.threshold_db if threshold_db is None else threshold_db)
        gate = self._gate(R, c.min_range if min_range is None else min_range, c.max_range if max_range is None else max_range)
        det = self._alloc_detections(F, R, D, A, threshold_db, workspace)
        self._call("rs_detect", rds.data_ptr(), gate.data_ptr(), thr, c.det_eps, det.key.data_ptr(),
                   det.power.data_ptr(), det.flags.data_ptr(), det.lead.data_ptr(), det.count.data_ptr(),
                   det.nlead.data_ptr(), det.overflow.data_ptr(), det.nnear.data_ptr(), det.psum.data_ptr(),
                   det.seg_cap, F, R, D, A, self.stream)
        return det

    def range_doppler_detect(self, cube: torch.Tensor, out: Optional[torch.Tensor] = None,
                             threshold_db: Optional[float] = None, min_range: Optional[float] = None,
                             max_range: Optional[float] = None, workspace=False, defer_power: bool = False):
        """range_doppler() + detect() as one call: on 256 x 128 planes the detection runs inside the 2-D FFT kernel on the
        plane it holds on chip (rs_range_doppler_detect); other shapes run the two stages.  -> (rds, detections).
        defer_power: do not write `power` here (the fused kernel keeps no |X|^2 and nothing on the path to the velocity
        reads it): Detections.materialize_power() gathers it from the RDS on demand, angles(write_power=True) writes it
        from the snapshots it holds in registers anyway."""
        assert cube.is_cuda and cube.dtype == torch.complex64 and cube.dim() == 4 and cube.is_contiguous()
        F, A, C, S = cube.shape
        c = self.cfg
        tab, tw_s, tw_c = self._fft_tables(S, C)[:3]
        thr = tables.power_threshold(c.threshold_db if threshold_db is None else threshold_db)
        gate = self._gate(S, c.min_range if min_range is None else min_range, c.max_range if max_range is None else max_range)
        rds = out if out is not None else torch.empty((F, S, A, C), dtype=torch.complex64, device=self.device)
        det = self._alloc_detections(F, S, C, A, threshold_db, workspace)
        fused_ok = S == 256 and C == 128 and A % 8 == 0
        det.power_pending = bool(defer_power and fused_ok)
        mid = None if fused_ok else self._buf("mid", (F, S, A, C), torch.complex64)
        fws = self._buf("fused_ws", (int(self.lib.rs_fused_detect_ws_bytes(F, A)),), torch.uint8) if fused_ok else None
        self._call("rs_range_doppler_detect", cube.data_ptr(), tab.data_ptr(), tw_s.data_ptr(), tw_c.data_ptr(),
                   _lib.ptr(mid), rds.data_ptr(), _lib.ptr(fws), F, A, C, 0, C, S, int(c.dc_removal), gate.data_ptr(), thr,
                   c.det_eps, det.key.data_ptr(), 0 if det.power_pending else det.power.data_ptr(), det.flags.data_ptr(),
                   det.lead.data_ptr(), det.count.data_ptr(), det.nlead.data_ptr(), det.overflow.data_ptr(),
                   det.nnear.data_ptr(), det.psum.data_ptr(), det.seg_cap, self.stream)
        if det.power_pending:
            det.power_fill = lambda: self._call(
                "rs_detection_power", rds.data_ptr(), det.key.data_ptr(), det.lead.data_ptr(), det.nlead.data_ptr(),
                det.power.data_ptr(), det.seg_cap, det.ntiles, det.F, det.R, det.D, det.A, self.stream)
        return rds, det

    def angles(self, rds: torch.Tensor, det: Detections, method: Optional[str] = None,
               fuse_ls: bool = True, write_power: bool = False) -> Detections:
        c = self.cfg
        method = method or c.method
        if method not in _lib.METHODS:
            raise ValueError(f"Unknown method: {method}")
        t = self._angle_tables(det.A)
        esprit_scale = c.lambda_c / (2 * np.pi * c.spacing)                       # angle_estimation.py:218
        det.method = method
        fuse = fuse_ls and method != "esprit" and det.A <= 16
        det.ls_partials = self._buf("ls_partials" + det.tag, (det.F * det.ntiles, 8), torch.float64) if fuse else None
        self._call(
            "rs_angles",
            rds.data_ptr(), t["scan"].data_ptr(), t["stride"], _lib.ptr(t["steer64"]), t["grid_f32"].data_ptr(), t["G"],
            _lib.METHODS[method], c.tie_eps, esprit_scale, det.key.data_ptr(), det.lead.data_ptr(),
            det.nlead.data_ptr(), det.flags.data_ptr(), det.aidx.data_ptr(), det.adeg.data_ptr(), det.phase.data_ptr(),
            det.seg_cap, det.ntiles, det.F, det.R, det.D, det.A,
            t["grid_cs"].data_ptr(), _lib.ptr(det.ls_partials), int(t["symmetric"]), det.ntie.data_ptr(), det.tielist.data_ptr(),
            _lib.ptr(t["mma"]), t["mma_tiles"],
            self._buf("cell_ws" + det.tag, (17 * det.F * det.R * det.D + 16,), torch.uint8).data_ptr() if det.A > 16 else
            (self._buf("pair_ws" + det.tag, (_lib.RS_ANGLES_WS_BYTES,), torch.uint8).data_ptr() if det.A == 16 else 0),
            _lib.ptr(t["tc"]), t["tc_halves"], det.power.data_ptr() if (write_power and det.power_pending) else 0,
            _lib.ptr(det.nnear), self.stream)
        if write_power:
            det.power_pending = False
        return det

    def velocity(self, det: Detections, out: Optional[torch.Tensor] = None, lambda_c: Optional[float] = None,
                 dt: Optional[float] = None, use_grid: bool = True) -> torch.Tensor:
        """-> float64 [F, 8] = v_x v_y v_z w_x w_y w_z success n_targets."""
        c = self.cfg
        lam = lambda_c or c.lambda_c_solver or c.lambda_c
        k = 4 * np.pi * (c.dt if dt is None else dt) / lam                        # velocity_solver.py:109
        t = self._angle_tables(det.A)
        vel = out if out is not None else torch.empty((det.F, 8), dtype=torch.float64, device=self.device)
        assert vel.is_contiguous() and vel.dtype == torch.float64
        if det.ls_partials is not None and use_grid and c.irls_iters == 0:
            self._call("rs_velocity_from_partials", det.ls_partials.data_ptr(), det.ntiles, det.F, k, c.velocity_bound,
                       det.overflow.data_ptr(), vel.data_ptr(), self.stream)
            return vel
        if c.irls_iters == 0:
            # sums per segment on all SMs, then the same per-frame reduction + solve as the fused path
            part = self._buf("ls_partials_lists" + det.tag, (det.F * det.ntiles, 8), torch.float64)
            self._call("rs_velocity_partials", det.aidx.data_ptr(), det.adeg.data_ptr(), det.phase.data_ptr(),
                       det.flags.data_ptr(), det.count.data_ptr(), t["grid_cs"].data_ptr() if use_grid else 0,
                       part.data_ptr(), det.seg_cap, det.ntiles, det.F, self.stream)
            self._call("rs_velocity_from_partials", part.data_ptr(), det.ntiles, det.F, k, c.velocity_bound,
                       det.overflow.data_ptr(), vel.data_ptr(), self.stream)
            return vel
        self._call(
            "rs_velocity_ls",
            det.aidx.data_ptr(), det.adeg.data_ptr(), det.phase.data_ptr(), det.flags.data_ptr(), det.count.data_ptr(),
            t["grid_cs"].data_ptr() if use_grid else 0, k, c.velocity_bound, c.irls_iters, c.huber_delta,
            vel.data_ptr(), det.seg_cap, det.ntiles, det.F, self.stream)
        return vel

    # ------------------------------------------------------------------ general-covariance MUSIC (Jacobi)
    def music_covariance(self, cov: torch.Tensor, num_sources: int = 1, sweeps: int = 10, want_vectors: bool = False):
        """MUSIC for arbitrary Hermitian covariances complex64 [n, A, A] (2 <= A <= 32): Jacobi eigendecomposition in
        registers, noise subspace V[:, num_sources:], pseudo-spectrum over this pipeline's azimuth grid
        (angle_estimation.py:109-154 generalised beyond the rank-1 single-snapshot case).
        Returns dict(eigvals [n,A] descending, spectrum [n,G], aidx [n], angle_deg [n], eigvecs [n,A,A] optional)."""
        assert cov.is_cuda and cov.dtype == torch.complex64 and cov.dim() == 3 and cov.shape[1] == cov.shape[2]
        cov = cov.contiguous()
        n, A, _ = cov.shape
        c = self.cfg
        key = ("steer64_any", A)
        if key not in self._tab:
            grid = tables.azimuth_grid(c.search_range, c.search_resolution)
            self._tab[key] = (self._dev(tables.steering(grid, np.arange(A) * c.spacing, c.lambda_c).astype(np.complex64)),
                              self._dev(grid.astype(np.float32)))
        steer, grid = self._tab[key]
        G = steer.shape[1]
        vals = torch.empty((n, A), dtype=torch.float32, device=self.device)
        spec = torch.empty((n, G), dtype=torch.float32, device=self.device)
        aidx = torch.empty((n,), dtype=torch.int32, device=self.device)
        vecs = torch.empty((n, A, A), dtype=torch.complex64, device=self.device) if want_vectors else None
        self._call("rs_music_covariance", cov.data_ptr(), n, A, int(num_sources), steer.data_ptr(), G, int(sweeps),
                   vals.data_ptr(), _lib.ptr(vecs), spec.data_ptr(), aidx.data_ptr(), self.stream)
        out = {"eigvals": vals, "spectrum": spec, "aidx": aidx, "angle_deg": grid[aidx.long()]}
        if want_vectors:
            out["eigvecs"] = vecs
        return out

    # ------------------------------------------------------------------ fp64 recheck of flagged decisions
    def _stats_buf(self, name: str) -> torch.Tensor:
        return self._buf(name, (4,), torch.int32)

    def recheck_detections(self, cube: torch.Tensor, det: Detections,
                           chirp_subset: Optional[Tuple[int, int]] = None) -> torch.Tensor:
        """Exact (fp64, from the raw cube) local-maximum / threshold decision for every RS_FLAG_NEARMAX entry.
        Before angles() it only edits the flags; after angles() (det.method set) it also moves a detection that
        changes state into / out of the fused velocity sums.  Must precede recheck_angles().
        Returns the device stats int32 [4] = rechecked, dropped, promoted, unresolved."""
        F, A, C, S = cube.shape
        c0, c1 = (0, C) if chirp_subset is None else chirp_subset
        _, _, _, tab128, tws128, twc128 = self._fft_tables(S, c1 - c0)
        stats = self._stats_buf("recheck_det_stats" + det.tag)
        post = det.method is not None and det.ls_partials is not None
        t = self._angle_tables(A)
        self._call("rs_recheck_detections_f64", cube.data_ptr(), tab128.data_ptr(), tws128.data_ptr(), twc128.data_ptr(),
                   C, c0, int(self.cfg.dc_removal),
                   float(10.0 ** (det.threshold_db / 10.0)), det.key.data_ptr(), det.flags.data_ptr(),
                   det.count.data_ptr(), det.nnear.data_ptr(), det.seg_cap, det.ntiles, F, A, c1 - c0, S,
                   det.aidx.data_ptr() if post else 0, det.phase.data_ptr() if post else 0,
                   t["grid_cs"].data_ptr() if post else 0, det.ls_partials.data_ptr() if post else 0,
                   stats.data_ptr(), self.stream)
        return stats

    def recheck_angles(self, cube: torch.Tensor, rds: torch.Tensor, det: Detections,
                       chirp_subset: Optional[Tuple[int, int]] = None, exhaustive: bool = False) -> torch.Tensor:
        """Exact grid argmax for every RS_FLAG_TIE / RS_FLAG_GUARD cell (grid methods).  Run after angles().
        Returns the device stats int32 [4] = rechecked, index changed, needed the fp64 snapshot, unresolved."""
        method = det.method or self.cfg.method
        stats = self._stats_buf("recheck_ang_stats" + det.tag)
        if method == "esprit":
            stats.zero_()
            return stats
        F, A, C, S = cube.shape
        c0, c1 = (0, C) if chirp_subset is None else chirp_subset
        _, _, _, tab128, tws128, twc128 = self._fft_tables(S, c1 - c0)
        t = self._angle_tables(A)
        self._call("rs_recheck_angles_f64", cube.data_ptr(), tab128.data_ptr(), tws128.data_ptr(), twc128.data_ptr(),
                   C, c0, int(self.cfg.dc_removal),
                   rds.data_ptr(), t["steer128"].data_ptr(), t["grid_f32"].data_ptr(), t["grid_cs"].data_ptr(), t["G"],
                   _lib.METHODS[method], float(self.cfg.fft_eps), det.psum.data_ptr(), det.ntie.data_ptr(),
                   det.tielist.data_ptr(), det.key.data_ptr(), det.lead.data_ptr(),
                   det.nlead.data_ptr(), det.flags.data_ptr(), det.aidx.data_ptr(), det.adeg.data_ptr(),
                   det.phase.data_ptr(), _lib.ptr(det.ls_partials), det.seg_cap, det.ntiles, F, A, c1 - c0, S,
                   self._buf("rc_idx" + det.tag, (F * det.ntiles * 16,), torch.int32).data_ptr(),
                   self._buf("rc_cnt" + det.tag, (F * det.ntiles,), torch.int32).data_ptr(),
                   self._buf("rc_snap" + det.tag, (F * det.ntiles * 16 * A,), torch.complex128).data_ptr(),
                   self._buf("rc_fcnt" + det.tag, (F,), torch.int32).data_ptr(),
                   self._buf("rc_flist" + det.tag, (F * _lib.RS_RECHECK_FRAME_CAP * 4,), torch.int32).data_ptr(),
                   stats.data_ptr(), self.stream)
        if exhaustive:
            # one pass settles at most 16 undecided cells per segment / 256 per frame; adversarial inputs (e.g. a
            # noise-free frame where every cell sits on the MUSIC guard) need more passes -- each syncs on the stats
            total = stats.clone()
            for _ in range(256):
                if int(stats[3].item()) == 0:
                    break
                self.recheck_angles(cube, rds, det, chirp_subset, exhaustive=False)
                total[:3] += stats[:3]
                total[3] = stats[3]
            stats.copy_(total)
        return stats

    # ------------------------------------------------------------------ whole path
    def process(self, cube: torch.Tensor, chunk_frames: int = 512, vel_out: Optional[torch.Tensor] = None,
                keep: bool = False, join: bool = True, after_solve=None):
        """Device-resident batch: returns vel float64 [F, 8] (and the last chunk's RDS / detections when
        keep=True and the batch is a single chunk).
        join=False: return without making the caller's stream wait for the side stream that runs the fp64 recheck and
        the solve of the last chunk -- the next process() call then overlaps its FFT / angle kernels with them (the two
        workspace sets alternate across calls).  `vel` and `cube` must not be touched before join().
        after_solve: callable enqueued right behind the solve of the last chunk, on the stream that runs it (e.g. the
        all-gather of the velocity rows), so that it needs no join either."""
        F = cube.shape[0]
        vel = vel_out if vel_out is not None else torch.empty((F, 8), dtype=torch.float64, device=self.device)
        last = None
        # The fp64 recheck and the solve of chunk i run on a side stream while the main stream already works on
        # chunk i+1 (two workspace sets): the recheck kernels are latency bound on a handful of CTAs and hide
        # behind the bandwidth / issue bound main kernels.
        main = torch.cuda.current_stream(self.device)
        if "side_stream" not in self._ws:
            prio = int(os.environ.get("RS_SIDE_PRIORITY", "-1"))     # high priority: its few CTAs slot in between the main kernels'
            self._ws["side_stream"] = torch.cuda.Stream(self.device, priority=prio)
        side = self._ws["side_stream"]
        side_done = self._ws.setdefault("side_done", [None, None])
        overlap = self.cfg.recheck and not keep and os.environ.get("RS_NO_OVERLAP") != "1"
        fuse_detect = os.environ.get("RS_SPLIT_DETECT") != "1"      # 1: rs_range_doppler_fft and rs_detect as two calls
        if not overlap and self._pending is not None:
            self._enqueue_recheck(None)

        for lo in range(0, F, chunk_frames):
            hi = min(F, lo + chunk_frames)
            n = hi - lo
            _, A, C, S = cube.shape
            ci = self._chunk_no                                    # alternates across calls too (join=False)
            self._chunk_no += 1
            tag = str(ci & 1)
            if side_done[ci & 1] is not None:
                main.wait_event(side_done[ci & 1])                 # workspace set free again
                side_done[ci & 1] = None
            rds_out = None if keep else self._buf("rds" + tag, (n, S, A, C), torch.complex64)
            if fuse_detect:
                rds, det = self.range_doppler_detect(cube[lo:hi], out=rds_out, workspace=False if keep else tag,
                                                     defer_power=True)
            else:
                rds = self.range_doppler(cube[lo:hi], out=rds_out)
                det = self.detect(rds, workspace=False if keep else tag)
            if overlap and self._pending is not None:       # RS_RECHECK_LATE=1 (see below)
                after = torch.cuda.Event()
                after.record(main)
                self._enqueue_recheck(after)
            self.angles(rds, det)
            if overlap:
                ready = torch.cuda.Event()
                ready.record(main)
                self._pending = (cube[lo:hi], vel[lo:hi], rds, det, ready, ci & 1, after_solve if hi == F else None)
                # The recheck + solve of this chunk go to the side stream right away: they run beside the NEXT chunk's 2-D
                # FFT, whose persistent clusters leave 16 SMs to other work (3.15 ms per 1000-frame step against 3.44 with the
                # recheck in line; profiles/overlap_probe.py, profiles/side_share_probe.sh).  Beside the angle scan instead
                # (RS_RECHECK_LATE=1) nothing is gained: the scan needs all 8 of its CTAs per SM resident, and every CTA the
                # recheck takes slows it by as much as the recheck gains (1.27 -> 1.97 ms while 0.73 ms of recheck runs,
                # profiles/timeline_probe.py).
                if os.environ.get("RS_RECHECK_LATE") != "1":
                    self._enqueue_recheck(None)
            else:
                if self.cfg.recheck:
                    self.recheck_detections(cube[lo:hi], det)
                    self.recheck_angles(cube[lo:hi], rds, det, exhaustive=keep)
                self.velocity(det, out=vel[lo:hi])
                if after_solve is not None and hi == F:
                    after_solve()
            last = (rds, det)
        if join:
            self.join()
        return (vel, last[0], last[1]) if keep else vel

    def _enqueue_recheck(self, after: Optional[torch.cuda.Event]) -> None:
        """fp64 recheck + solve of the pending chunk on the side stream (after its angle scan, and after `after`)."""
        jcube, jvel, jrds, jdet, jready, jslot, jafter = self._pending
        self._pending = None
        side = self._ws["side_stream"]
        with torch.cuda.stream(side):
            side.wait_event(jready)
            if after is not None:
                side.wait_event(after)
            self.recheck_detections(jcube, jdet)
            self.recheck_angles(jcube, jrds, jdet)
            self.velocity(jdet, out=jvel)
            if jafter is not None:
                jafter()
            done = torch.cuda.Event()
            done.record(side)
            self._ws["side_done"][jslot] = done

    def join(self) -> None:
        """Enqueue what process(join=False) left pending and make the current stream wait for the side stream."""
        if self._pending is not None:
            self._enqueue_recheck(None)
        side = self._ws.get("side_stream")
        if side is not None:
            torch.cuda.current_stream(self.device).wait_stream(side)

    def status(self) -> Dict[str, int]:
        """What the last process() call could not settle (synchronises): frames whose detection segments overflowed --
        their velocity row reports success = 0 -- and flagged decisions the fp64 recheck left at their fp32 value (one
        recheck pass settles 16 cells per segment / 256 per frame; process(keep=True) iterates until none is left).
        Counters of the chunks that used the two workspace sets last."""
        torch.cuda.synchronize(self.device)
        out = {"overflow_frames": 0, "unresolved_detections": 0, "unresolved_angles": 0}
        for tag in ("", "0", "1"):
            ov = self._ws.get("det_overflow" + tag)
            if ov is not None:
                out["overflow_frames"] += int((ov[: self._ws.get("_nframes" + tag, 0)] != 0).sum().item())
            for key, name in (("recheck_det_stats", "unresolved_detections"), ("recheck_ang_stats", "unresolved_angles")):
                st = self._ws.get(key + tag)
                if st is not None:
                    out[name] += int(st[3].item())
        return out

    def process_host(self, cube_host: torch.Tensor, chunk_frames: int = 32, vel_dev: Optional[torch.Tensor] = None,
                     vel_host: Optional[torch.Tensor] = None, copy_only: bool = False) -> torch.Tensor:
        """End-to-end with HOST buffers: pinned cube[F,A,C,S] -> pinned vel[F,8]; H2D and D2H copies run on
        a second stream and overlap the kernels of the neighbouring chunk.
        vel_dev: device rows to write (e.g. this rank's slot of the all-gather buffer); vel_host: pinned result buffer;
        copy_only: the same buffers, chunking and stream protocol with the kernels left out -- the platform's ceiling
        for this path (bench.py reports the end-to-end rate as a fraction of it)."""
        assert not cube_host.is_cuda and cube_host.dtype == torch.complex64
        F, A, C, S = cube_host.shape
        if vel_dev is None:
            vel_dev = self._buf("host_vel_dev", (F, 8), torch.float64)
        if vel_host is None:
            vel_host = torch.empty((F, 8), dtype=torch.float64, pin_memory=True)
        copy_stream = self._ws.setdefault("copy_stream", torch.cuda.Stream(self.device))
        main = torch.cuda.current_stream(self.device)
        stage = [self._buf(f"stage{i}", (chunk_frames, A, C, S), torch.complex64) for i in range(2)]
        ev_h2d = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        for ci, lo in enumerate(range(0, F, chunk_frames)):
            hi = min(F, lo + chunk_frames)
            b = ci & 1
            with torch.cuda.stream(copy_stream):
                if ci >= 2:
                    copy_stream.wait_event(ev_free[b])
                stage[b][: hi - lo].copy_(cube_host[lo:hi], non_blocking=True)
                ev_h2d[b].record(copy_stream)
            main.wait_event(ev_h2d[b])
            if not copy_only:
                self.process(stage[b][: hi - lo], chunk_frames=chunk_frames, vel_out=vel_dev[lo:hi])
            ev_free[b].record(main)
        vel_host.copy_(vel_dev[:F], non_blocking=True)
        main.synchronize()
        return vel_host
