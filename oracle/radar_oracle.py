"""TEST INFRASTRUCTURE ONLY -- fp64 numpy restatement of radar-slam's per-frame hot path.

This is the checker for the CUDA path and the reported CPU baseline; it is never
imported by the product (see oracle/__init__.py).  Every function cites the
reference lines it follows (paths relative to /root/reference).  The restatement
is vectorised (the reference loops in Python over chirps / peaks / grid angles)
but performs the same floating-point operations in the same order wherever that
changes the result; where it cannot (LAPACK null-space bases, DE's RNG) the
docstring says so.

Pinned by tests/test_oracle_vs_reference.py (live reference, this container only)
and tests/test_oracle_golden.py (committed fixtures written by oracle/make_golden.py
from the live reference).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from collections import deque
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
from scipy.linalg import eigh as _sp_eigh, svd as _sp_svd
from scipy.ndimage import maximum_filter as _maximum_filter
from scipy.signal import windows as _windows

C0 = 3e8  # dechirp.py:61, angle_estimation.py:48, velocity_solver.py:50


# ----------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------
@dataclass
class RadarParams:
    """Constructor arguments of SignalPreprocessor / FMCWRadarSimulator and their derived
    constants (dechirp.py:29-68, simulate_raw.py:35-79)."""
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
    noise_power: float = 0.01

    @property
    def lambda_c(self) -> float:
        return C0 / self.fc

    @property
    def samples_per_chirp(self) -> int:
        return int(self.chirp_duration * self.sampling_rate)      # dechirp.py:63

    @property
    def chirp_rate(self) -> float:
        return self.bandwidth / self.chirp_duration               # dechirp.py:64

    @property
    def range_resolution(self) -> float:
        return C0 / (2 * self.bandwidth)                          # dechirp.py:67

    @property
    def spacing(self) -> float:
        return self.antenna_spacing or (self.lambda_c / 2)        # angle_estimation.py:50

    @property
    def antenna_positions(self) -> np.ndarray:
        return np.arange(self.num_antennas) * self.spacing        # angle_estimation.py:56


# ----------------------------------------------------------------------------------
# dechirp / window / DC / range-Doppler spectrum
# ----------------------------------------------------------------------------------
def reference_chirp(p: RadarParams) -> np.ndarray:
    """dechirp.py:74-83.  linspace includes the endpoint; the phase reaches ~2e6 cycles so
    this must be evaluated in fp64."""
    t = np.linspace(0, p.chirp_duration, p.samples_per_chirp)
    phase = 2 * np.pi * (p.fc * t + 0.5 * p.chirp_rate * t ** 2)
    return np.exp(1j * phase)


def window(window_type: str, n: int) -> np.ndarray:
    """dechirp.py:96-106 -- scipy *symmetric* windows; ValueError on an unknown name."""
    if window_type == "hann":
        return _windows.hann(n)
    if window_type == "hamming":
        return _windows.hamming(n)
    if window_type == "blackman":
        return _windows.blackman(n)
    raise ValueError(f"Unknown window type: {window_type}")


def range_doppler_spectrum(frame: np.ndarray, p: RadarParams,
                           chirp_subset: Optional[Tuple[int, int]] = None) -> np.ndarray:
    """dechirp.py:168-213.  frame [A, C, S] complex -> RDS [A, S, C'] complex128.

    Per chirp the reference does (x * conj(ref)) * w - mean (dechirp.py:139,108,120, in that
    order, :156-164), stores it transposed (:205), then fft2 over (range, doppler) and fftshift
    over both axes (:208-211).  The same elementwise expressions are applied here to the whole
    cube; the result is bit-identical to the reference loop (tests/test_oracle_vs_reference.py).
    """
    frame = np.asarray(frame)
    if chirp_subset is not None:                                  # dechirp.py:184-187
        a, b = chirp_subset
        frame = frame[:, a:b, :]
    ref = reference_chirp(p)
    w = window(p.window_type, frame.shape[2])
    x = frame * np.conj(ref)
    x = x * w
    if p.dc_removal:
        x = x - np.mean(x, axis=2, keepdims=True)
    rds = np.zeros((frame.shape[0], frame.shape[2], frame.shape[1]), dtype=complex)
    rds[:] = np.transpose(x, (0, 2, 1))
    out = np.fft.fft2(rds, axes=(1, 2))
    return np.fft.fftshift(out, axes=(1, 2))


def range_axis(p: RadarParams, range_bins: int) -> np.ndarray:
    return np.linspace(0, p.range_resolution * range_bins, range_bins)      # dechirp.py:241


def doppler_axis(p: RadarParams, doppler_bins: int) -> np.ndarray:
    return np.linspace(-p.sampling_rate / 2, p.sampling_rate / 2, doppler_bins)  # dechirp.py:242


def extract_peaks(rds: np.ndarray, p: RadarParams, threshold_db: float = -20.0,
                  min_range: float = 1.0, max_range: float = 200.0) -> Dict[str, np.ndarray]:
    """dechirp.py:215-278 as arrays instead of a list of dicts.

    power_db = 10 log10(|X|^2 + 1e-12) (:235-238); per antenna a 3x3 maximum_filter (scipy
    default mode 'reflect') equality test AND a strict '> threshold_db' (:250-254); np.where
    row-major order (:257); inclusive range gate on the linspace axis (:263).  Output order is
    antenna -> range_bin -> doppler_bin (:246-258)."""
    power_db = 10 * np.log10(np.abs(rds) ** 2 + 1e-12)
    r_axis = range_axis(p, rds.shape[1])
    d_axis = doppler_axis(p, rds.shape[2])
    local_max = _maximum_filter(power_db, size=(1, 3, 3)) == power_db
    mask = local_max & (power_db > threshold_db)
    gate = (r_axis >= min_range) & (r_axis <= max_range)
    mask &= gate[None, :, None]
    a, r, d = np.nonzero(mask)                     # C order == antenna, range, doppler
    return {
        "antenna": a.astype(np.int64), "range_bin": r.astype(np.int64), "doppler_bin": d.astype(np.int64),
        "range_m": r_axis[r], "doppler_hz": d_axis[d], "power_db": power_db[a, r, d],
        "range_bins_m": r_axis, "doppler_bins_hz": d_axis, "power_spectrum_db": power_db,
    }


def peaks_as_dicts(pk: Dict[str, np.ndarray]) -> List[dict]:
    """The list-of-dict form the reference returns in peak_info['peaks'] (dechirp.py:264-271)."""
    return [{"antenna": int(a), "range_bin": int(r), "doppler_bin": int(d), "range_m": rm,
             "doppler_hz": dh, "power_db": pw}
            for a, r, d, rm, dh, pw in zip(pk["antenna"], pk["range_bin"], pk["doppler_bin"],
                                           pk["range_m"], pk["doppler_hz"], pk["power_db"])]


# ----------------------------------------------------------------------------------
# angle estimation
# ----------------------------------------------------------------------------------
def azimuth_grid(search_range=(-90, 90), search_resolution: float = 0.5) -> np.ndarray:
    return np.arange(search_range[0], search_range[1] + search_resolution, search_resolution)  # angle_estimation.py:59


def steering_matrix(grid_deg: np.ndarray, positions: np.ndarray, lambda_c: float) -> np.ndarray:
    """angle_estimation.py:102-107 for every grid angle: [G, M] complex128."""
    az = np.radians(grid_deg)
    phases = 2 * np.pi * positions[None, :] * np.sin(az)[:, None] / lambda_c
    return np.exp(1j * phases)


def spatial_signatures(rds: np.ndarray, range_bin: np.ndarray, doppler_bin: np.ndarray) -> np.ndarray:
    """angle_estimation.py:83-88 for a batch of cells: rds[:, r, d] / sqrt(sum |.|^2) if > 0.  [D, M]."""
    s = rds[:, range_bin, doppler_bin].T.copy()
    power = np.sum(np.abs(s) ** 2, axis=1)
    nz = power > 0
    s[nz] = s[nz] / np.sqrt(power[nz])[:, None]
    return s


def music_spectrum_literal(sig: np.ndarray, grid_deg: np.ndarray, positions: np.ndarray,
                           lambda_c: float, num_sources: int = 1) -> np.ndarray:
    """angle_estimation.py:127-152 for ONE signature, same call sequence (scipy eigh of the
    outer product, descending sort, noise subspace, per-angle 4-matrix product, 1e-12 guard)."""
    R = np.outer(sig, sig.conj())
    vals, vecs = _sp_eigh(R)
    idx = np.argsort(vals)[::-1]
    vecs = vecs[:, idx]
    En = vecs[:, num_sources:]
    out = np.zeros(len(grid_deg))
    for i, az in enumerate(grid_deg):
        a = np.exp(1j * (2 * np.pi * positions * np.sin(np.radians(az)) / lambda_c))
        den = np.abs(a.conj().T @ En @ En.conj().T @ a)
        out[i] = 1.0 / den if den > 1e-12 else 0.0
    return out


def music_spectra(sigs: np.ndarray, steering: np.ndarray, num_sources: int = 1,
                  chunk: int = 1024) -> np.ndarray:
    """Batched angle_estimation.py:127-152: eigh of every rank-1 covariance, noise subspace
    E_n = V[:, num_sources:], denominator a^H E_n E_n^H a, 1e-12 guard.  [D, G] float64.
    The null-space basis LAPACK returns is arbitrary but the projector E_n E_n^H is not; values
    agree with the literal loop to ~1e-15 of M (for num_sources == 1)."""
    D, M = sigs.shape
    out = np.empty((D, steering.shape[0]))
    Ac = steering.conj()                                                   # [G, M]
    for lo in range(0, D, chunk):
        s = sigs[lo:lo + chunk]
        R = s[:, :, None] * s.conj()[:, None, :]
        vals, vecs = np.linalg.eigh(R)                                     # ascending
        En = vecs[:, :, ::-1][:, :, num_sources:]                          # [d, M, M-K]
        T = np.einsum("gm,dmj->dgj", Ac, En, optimize=True)
        den = np.abs(np.sum(T * T.conj(), axis=2))
        with np.errstate(divide="ignore"):
            out[lo:lo + chunk] = np.where(den > 1e-12, 1.0 / den, 0.0)
    return out


def beamforming_spectra(sigs: np.ndarray, steering: np.ndarray) -> np.ndarray:
    """angle_estimation.py:239-245 / robust_angle_estimation.py:237-241: |a^H s|^2.  [D, G]."""
    return np.abs(sigs @ steering.conj().T) ** 2


def argmax_angles(spectra: np.ndarray, grid_deg: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """angle_estimation.py:173-174 -- np.argmax returns the FIRST maximum."""
    idx = np.argmax(spectra, axis=1)
    return idx, grid_deg[idx]


def esprit_angle_literal(sig: np.ndarray, lambda_c: float, spacing: float, num_sources: int = 1) -> float:
    """angle_estimation.py:195-225 for ONE signature with the same library calls."""
    try:
        U, s, Vh = _sp_svd(np.column_stack([sig[:-1], sig[1:]]))
        Us = U[:, :num_sources]
        Phi = np.linalg.pinv(Us[:-1, :]) @ Us[1:, :]
        ev = np.linalg.eigvals(Phi)
        phase = np.angle(ev[0])
        return float(np.degrees(np.arcsin(phase * lambda_c / (2 * np.pi * spacing))))
    except Exception:
        return 0.0


def esprit_angles(sigs: np.ndarray, lambda_c: float, spacing: float) -> np.ndarray:
    """Closed form of angle_estimation.py:195-221 for num_sources == 1.

    The first left singular vector of B = [s[:-1], s[1:]] is u ~ B v with v the principal
    eigenvector of the 2x2 Hermitian B^H B; Phi = (u[:-1]^H u[1:]) / (u[:-1]^H u[:-1]) is
    invariant to the scale/phase of u, so theta = asin(arg(u[:-1]^H u[1:]) lambda / (2 pi d)).
    Matches the literal routine to ~1e-12 deg (tests/test_oracle_vs_reference.py)."""
    b1, b2 = sigs[:, :-1], sigs[:, 1:]
    alpha = np.sum(np.abs(b1) ** 2, axis=1)
    gamma = np.sum(np.abs(b2) ** 2, axis=1)
    beta = np.sum(b1.conj() * b2, axis=1)
    # principal eigenvector of [[alpha, beta], [conj(beta), gamma]]
    half = 0.5 * (alpha - gamma)
    lam = 0.5 * (alpha + gamma) + np.sqrt(half ** 2 + np.abs(beta) ** 2)
    v0a, v1a = beta, lam - alpha                 # (A - lam I) v = 0, row 1
    v0b, v1b = lam - gamma, beta.conj()          # row 2
    use_a = (np.abs(v0a) ** 2 + np.abs(v1a) ** 2) >= (np.abs(v0b) ** 2 + np.abs(v1b) ** 2)
    v0 = np.where(use_a, v0a, v0b)
    v1 = np.where(use_a, v1a, v1b)
    u = b1 * v0[:, None] + b2 * v1[:, None]
    num = np.sum(u[:, :-1].conj() * u[:, 1:], axis=1)
    phase = np.angle(num)
    with np.errstate(invalid="ignore"):
        return np.degrees(np.arcsin(phase * lambda_c / (2 * np.pi * spacing)))


# ----------------------------------------------------------------------------------
# robust angle estimation (stateful)
# ----------------------------------------------------------------------------------
def angle_confidence(sig: np.ndarray, angle_deg: float, positions: np.ndarray, lambda_c: float) -> float:
    """robust_angle_estimation.py:101-138 for one signature."""
    a = np.exp(1j * (2 * np.pi * positions * np.sin(np.radians(angle_deg)) / lambda_c))
    corr = np.abs(a.conj().T @ sig)
    sp = np.sum(np.abs(sig) ** 2)
    ncorr = corr / np.sqrt(sp) if sp > 0 else 0.0
    perr = np.mean(np.abs(np.angle(np.exp(1j * (np.angle(sig) - np.angle(a))))))
    pcons = np.exp(-perr)
    pw = np.abs(sig) ** 2
    nf = np.percentile(pw, 20)
    if nf > 0:
        snr_c = min(1.0, np.log10(np.mean(pw) / nf) / 3.0)
    else:
        snr_c = 0.0
    c = ncorr * 0.4 + pcons * 0.3 + snr_c * 0.3
    return min(1.0, max(0.0, c))


def multipath_analysis(sig: np.ndarray) -> dict:
    """robust_angle_estimation.py:151-218 for one signature (MDL on the rank-1 eigenvalues).
    Because the 'geometric' and arithmetic noise means are the same expression (:177-179) the
    log term vanishes and the penalty grows with k, so num_sources is 1 whenever the k=1 noise
    mean is positive."""
    R = np.outer(sig, sig.conj())
    vals, _ = _sp_eigh(R)
    vals = np.real(vals)
    vals = vals[np.argsort(vals)[::-1]]
    N = len(vals)
    mdl = []
    for k in range(1, min(N, 5)):
        L = N - k
        noise = vals[k:]
        if L > 0 and len(noise) > 0 and np.mean(noise) > 0:
            nm = np.mean(noise)
            mdl.append(-L * np.log(nm / nm) + 0.5 * k * (2 * N - k) * np.log(L))
        else:
            mdl.append(float("inf"))
    if mdl and min(mdl) != float("inf"):
        ns = int(np.argmin(mdl)) + 1
    else:
        ns = 1
    if N > 1:
        sp, npow = np.sum(vals[:ns]), np.sum(vals[ns:])
        snr = sp / npow if npow > 0 else float("inf")
        cond = vals[0] / vals[-1] if vals[-1] > 0 else float("inf")
    else:
        snr = cond = float("inf")
    return {"num_sources": ns, "snr_ratio": snr, "condition_number": cond, "eigenvalues": vals,
            "is_multipath": ns > 1,
            "interference_level": min(1.0, 1.0 / snr) if snr > 0 else 1.0}


class RobustOracle:
    """robust_angle_estimation.py:23-411 -- power filter, stable sort, top-K, beamforming argmax,
    confidence, temporal smoothing (sequential host state), reliability gate."""

    def __init__(self, p: RadarParams, search_range=(-90, 90), search_resolution=1.0, temporal_window=5,
                 confidence_threshold=0.7, smoothing_factor=0.3, max_targets=100):
        self.p = p
        self.grid = azimuth_grid(search_range, search_resolution)
        self.positions = p.antenna_positions
        self.steering = steering_matrix(self.grid, self.positions, p.lambda_c)
        self.temporal_window = temporal_window
        self.confidence_threshold = confidence_threshold
        self.smoothing_factor = smoothing_factor
        self.max_targets = max_targets
        self.angle_history: dict = {}
        self.confidence_history: dict = {}

    def smooth(self, tid: str, angle: float, conf: float) -> Tuple[float, float]:
        """robust_angle_estimation.py:290-330."""
        if tid not in self.angle_history:
            self.angle_history[tid] = deque(maxlen=self.temporal_window)
            self.confidence_history[tid] = deque(maxlen=self.temporal_window)
        self.angle_history[tid].append(angle)
        self.confidence_history[tid].append(conf)
        if len(self.angle_history[tid]) >= 2:
            ang = np.array(self.angle_history[tid])
            cf = np.array(self.confidence_history[tid])
            w = cf / np.sum(cf) if np.sum(cf) > 0 else np.ones_like(cf) / len(cf)
            ar = np.radians(ang)
            sm = np.degrees(np.arctan2(np.sum(w * np.sin(ar)), np.sum(w * np.cos(ar))))
            prev = self.angle_history[tid][-2]
            sm = self.smoothing_factor * sm + (1 - self.smoothing_factor) * prev
            return sm, np.mean(cf)
        return angle, conf

    def select(self, pk: Dict[str, np.ndarray]) -> np.ndarray:
        """robust_angle_estimation.py:362-365: power_db > -25, stable descending sort, top-K.
        Returns indices into the peak arrays."""
        keep = np.nonzero(pk["power_db"] > -25.0)[0]
        order = np.argsort(-pk["power_db"][keep], kind="stable")
        return keep[order][: self.max_targets]

    def process(self, rds: np.ndarray, pk: Dict[str, np.ndarray], frame_timestamp=None) -> List[dict]:
        """robust_angle_estimation.py:346-411."""
        sel = self.select(pk)
        out = []
        for i in sel:
            r, d = int(pk["range_bin"][i]), int(pk["doppler_bin"][i])
            sig = rds[:, r, d]
            pw = np.sum(np.abs(sig) ** 2)
            if pw > 0:
                sig = sig / np.sqrt(pw)
            tid = f"target_{r}_{d}"
            interf = multipath_analysis(sig)
            spec = np.abs(self.steering.conj() @ sig) ** 2
            init = self.grid[int(np.argmax(spec))]
            conf = angle_confidence(sig, init, self.positions, self.p.lambda_c)
            ang, sconf = self.smooth(tid, init, conf)
            reliable = (sconf >= self.confidence_threshold) and (not interf["is_multipath"])
            if reliable:
                out.append({"range_m": pk["range_m"][i], "doppler_hz": pk["doppler_hz"][i],
                            "power_db": pk["power_db"][i], "azimuth_deg": ang, "azimuth_rad": np.radians(ang),
                            "confidence": sconf, "is_reliable": reliable, "interference_analysis": interf,
                            "antenna": int(pk["antenna"][i]), "range_bin": r, "doppler_bin": d,
                            "spatial_signature": sig, "target_id": tid, "initial_angle": init,
                            "raw_confidence": conf, "timestamp": frame_timestamp})
        return out


# ----------------------------------------------------------------------------------
# ego-velocity
# ----------------------------------------------------------------------------------
def observed_phases(sigs: np.ndarray) -> np.ndarray:
    """velocity_solver.py:136 -- inter-antenna phase angle(s[1] conj(s[0])) per target."""
    return np.angle(sigs[:, 1] * np.conj(sigs[:, 0]))


def phase_model(positions: np.ndarray, angles: np.ndarray, v: np.ndarray, w: np.ndarray,
                dt: float, lambda_c: float) -> np.ndarray:
    """velocity_solver.py:84-113 vectorised: 4 pi dt/lambda * d . (v + w x r)."""
    az, el = angles[:, 0], angles[:, 1]
    d = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], axis=1)
    rel = v[None, :] + np.cross(np.broadcast_to(w, positions.shape), positions)
    return (4 * np.pi * np.sum(rel * d, axis=1) * dt) / lambda_c


def box_ls_2d(c: np.ndarray, s: np.ndarray, y: np.ndarray, k: float, bound: float = 50.0,
              weights: Optional[np.ndarray] = None) -> Tuple[np.ndarray, bool]:
    """Exact minimiser of sum w (y - k (vx c + vy s))^2 over the box |vx|,|vy| <= bound: interior
    normal-equation solution if feasible, otherwise the best of the four clipped edge minimisers.
    Returns (v[2], well_conditioned)."""
    w = np.ones_like(y) if weights is None else weights
    g11, g22, g12 = k * k * np.sum(w * c * c), k * k * np.sum(w * s * s), k * k * np.sum(w * c * s)
    b1, b2 = k * np.sum(w * y * c), k * np.sum(w * y * s)

    def f(vx, vy):
        return g11 * vx * vx + 2 * g12 * vx * vy + g22 * vy * vy - 2 * (b1 * vx + b2 * vy)

    det = g11 * g22 - g12 * g12
    ok = det > 1e-12 * max(g11 * g22, 1e-300)
    if ok:
        vx = (g22 * b1 - g12 * b2) / det
        vy = (g11 * b2 - g12 * b1) / det
        if abs(vx) <= bound and abs(vy) <= bound:
            return np.array([vx, vy]), True
    cands = []
    for vx in (-bound, bound):
        vy = (b2 - g12 * vx) / g22 if g22 > 0 else 0.0
        cands.append((vx, min(bound, max(-bound, vy))))
    for vy in (-bound, bound):
        vx = (b1 - g12 * vy) / g11 if g11 > 0 else 0.0
        cands.append((min(bound, max(-bound, vx)), vy))
    if not ok:
        # rank-deficient (all directions parallel): minimum-norm solution, clipped
        tr = g11 + g22
        if tr > 0:
            cands.append((min(bound, max(-bound, b1 / tr)), min(bound, max(-bound, b2 / tr))))
        else:
            cands.append((0.0, 0.0))
    best = min(cands, key=lambda q: f(*q))
    return np.array(best), bool(ok)


def solve_velocity(range_m: np.ndarray, azimuth_rad: np.ndarray, sigs: np.ndarray, lambda_c: float,
                   dt: float = 0.1) -> dict:
    """velocity_solver.py:309-355 + :178-307 with the optimiser replaced by what it converges to.

    The reference builds pos = range [cos az, sin az, 0] and direction [cos az, sin az, 0]
    (:337-342, elevation hard-wired to 0), so (w x pos).dir == 0 and the v_z column is zero:
    the cost (:171-174) is a convex quadratic in (v_x, v_y) only.  differential_evolution(seed=42)
    with bounds +-50 (:216-222, :250-257) converges to the box-constrained least-squares point;
    v_z and omega are unobservable (DE returns RNG-dependent values for them) and are reported
    as 0 here.  N < 3 -> {'success': False, 'message': 'Insufficient targets'} (:202-204)."""
    N = len(range_m)
    if N < 3:
        return {"success": False, "message": "Insufficient targets"}
    y = observed_phases(sigs)
    k = 4 * np.pi * dt / lambda_c
    v2, well = box_ls_2d(np.cos(azimuth_rad), np.sin(azimuth_rad), y, k)
    v = np.array([v2[0], v2[1], 0.0])
    w = np.zeros(3)
    pos = np.stack([range_m * np.cos(azimuth_rad), range_m * np.sin(azimuth_rad), np.zeros(N)], axis=1)
    ang = np.stack([azimuth_rad, np.zeros(N)], axis=1)
    pred = phase_model(pos, ang, v, w, dt, lambda_c)
    res = y - pred
    return {"success": True, "velocity": v, "angular_velocity": w, "cost": float(np.sum(res ** 2)),
            "rmse": float(np.sqrt(np.mean(res ** 2))), "max_residual": float(np.max(np.abs(res))),
            "residuals": res, "predicted_phases": pred, "observed_phases": y, "num_targets": N,
            "well_conditioned": well}


# ----------------------------------------------------------------------------------
# input synthesis (the input oracle; outside the accelerated path)
# ----------------------------------------------------------------------------------
def scatterer_response(p: RadarParams, scatterers: np.ndarray) -> np.ndarray:
    """Noise-free part of simulate_raw.py:147-213: [A, S] complex128 (identical for every chirp,
    :190-209 never uses chirp_start_time).  scatterers rows: (range m, azimuth rad, rcs dB, vr)."""
    S = p.samples_per_chirp
    t = np.linspace(0, p.chirp_duration, S)
    ref = np.exp(1j * (2 * np.pi * (p.fc * t + 0.5 * p.chirp_rate * t ** 2)))
    pos = p.antenna_positions
    out = np.zeros((p.num_antennas, S), dtype=complex)
    for row in np.atleast_2d(scatterers):
        rng_m, az, rcs_db, vr = (float(x) for x in row[:4])
        if rng_m <= 0 or not np.isfinite([rng_m, az, rcs_db, vr]).all():   # simulate_raw.py:181
            continue
        delay = 2 * rng_m / C0
        amp = np.sqrt(10 ** (rcs_db / 10)) / (4 * np.pi * rng_m ** 2)
        dop = 4 * np.pi * vr * p.fc / C0
        aph = np.zeros(p.num_antennas, dtype=complex)
        for i in range(p.num_antennas):
            aph[i] = amp * np.exp(1j * (dop + 2 * np.pi * pos[i] * np.sin(az) / p.lambda_c))
        td = t - delay
        valid = (td >= 0) & (td <= p.chirp_duration)
        if np.any(valid):
            tv = td[valid]
            delayed = np.exp(1j * (2 * np.pi * (p.fc * tv + 0.5 * p.chirp_rate * tv ** 2)))
            base = delayed * np.conj(ref[valid])
            for i in range(p.num_antennas):
                out[i, valid] += aph[i] * base
    return out


def synthesize_frame(p: RadarParams, scatterers: np.ndarray, rng=None) -> np.ndarray:
    """simulate_raw.py:147-221: scatterer term broadcast over chirps plus complex Gaussian noise
    drawn as randn(shape) then randn(shape) from the GLOBAL legacy numpy RNG (:216-219) unless an
    np.random.RandomState is passed.  [A, C, S] complex128."""
    sig = scatterer_response(p, scatterers)
    out = np.zeros((p.num_antennas, p.num_chirps, p.samples_per_chirp), dtype=complex)
    out += sig[:, None, :]
    r = np.random if rng is None else rng
    noise = np.sqrt(p.noise_power) * (r.randn(*out.shape) + 1j * r.randn(*out.shape))
    out += noise
    return out


# ----------------------------------------------------------------------------------
# whole-frame driver (what bench.py's cpu_baseline times)
# ----------------------------------------------------------------------------------
def process_frame(frame: np.ndarray, p: RadarParams, method: str = "music", search_resolution: float = 0.5,
                  threshold_db: float = -20.0, lambda_c_solver: Optional[float] = None, dt: float = 0.1,
                  timings: Optional[dict] = None) -> dict:
    """One pass of the hot path on one frame: RDS -> peaks -> signatures -> angles -> velocity."""
    import time
    t0 = time.perf_counter()
    rds = range_doppler_spectrum(frame, p)
    t1 = time.perf_counter()
    pk = extract_peaks(rds, p, threshold_db)
    t2 = time.perf_counter()
    grid = azimuth_grid((-90, 90), search_resolution)
    sigs = spatial_signatures(rds, pk["range_bin"], pk["doppler_bin"])
    if method == "esprit":
        ang = esprit_angles(sigs, p.lambda_c, p.spacing)
        idx = None
    else:
        A = steering_matrix(grid, p.antenna_positions, p.lambda_c)
        spec = music_spectra(sigs, A) if method == "music" else beamforming_spectra(sigs, A)
        idx, ang = argmax_angles(spec, grid)
    t3 = time.perf_counter()
    vel = solve_velocity(pk["range_m"], np.radians(ang), sigs, lambda_c_solver or p.lambda_c, dt)
    t4 = time.perf_counter()
    if timings is not None:
        for k_, v_ in (("rds", t1 - t0), ("peaks", t2 - t1), ("angles", t3 - t2), ("velocity", t4 - t3)):
            timings[k_] = timings.get(k_, 0.0) + v_
    return {"rds": rds, "peaks": pk, "signatures": sigs, "angle_idx": idx, "angle_deg": ang, "velocity": vel}
