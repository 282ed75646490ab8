"""Host-side constant tables for the CUDA path, evaluated in fp64 with the reference's own
expressions and rounded once (the chirp phase reaches ~2e6 cycles and cannot be
evaluated in fp32 on the device).  Pure numpy; no oracle import."""
from __future__ import annotations

from typing import Tuple

import numpy as np
from scipy.signal import windows as _windows

C0 = 3e8


def window(window_type: str, n: int) -> np.ndarray:
    """Symmetric scipy windows, ValueError on unknown names (dechirp.py:96-106)."""
    if window_type == "hann":
        return _windows.hann(n)
    if window_type == "hamming":
        return _windows.hamming(n)
    if window_type == "blackman":
        return _windows.blackman(n)
    raise ValueError(f"Unknown window type: {window_type}")


def reference_chirp(fc: float, chirp_rate: float, chirp_duration: float, n: int) -> np.ndarray:
    """dechirp.py:74-83."""
    t = np.linspace(0, chirp_duration, n)
    phase = 2 * np.pi * (fc * t + 0.5 * chirp_rate * t ** 2)
    return np.exp(1j * phase)


def dechirp_table(fc, chirp_rate, chirp_duration, n, window_type) -> np.ndarray:
    """conj(ref) * window, the whole per-sample factor of process_chirp (dechirp.py:139,108). c128[n]."""
    return np.conj(reference_chirp(fc, chirp_rate, chirp_duration, n)) * window(window_type, n)


def twiddles(n: int) -> np.ndarray:
    """exp(-2 pi i k / n), k < n, complex64 rounded from fp64."""
    k = np.arange(n)
    return np.exp(-2j * np.pi * k / n).astype(np.complex64)


def twiddles128(n: int) -> np.ndarray:
    """exp(-2 pi i k / n), k < n, complex128 (the fp64 recheck's direct DFT)."""
    k = np.arange(n)
    return np.exp(-2j * np.pi * k / n)


def azimuth_grid(search_range: Tuple[float, float], search_resolution: float) -> np.ndarray:
    """angle_estimation.py:59."""
    return np.arange(search_range[0], search_range[1] + search_resolution, search_resolution)


def steering(grid_deg: np.ndarray, positions: np.ndarray, lambda_c: float) -> np.ndarray:
    """exp(i 2 pi pos sin(az) / lambda) for every (antenna, grid angle): c128 [A][G]
    (angle_estimation.py:102-107, transposed so a grid scan reads contiguous memory)."""
    az = np.radians(grid_deg)
    phases = 2 * np.pi * positions[:, None] * np.sin(az)[None, :] / lambda_c
    return np.exp(1j * phases)


def padded_antennas(A: int) -> int:
    return 2 if A <= 2 else 4 if A <= 4 else 8 if A <= 8 else 16


def scan_table(grid_deg: np.ndarray, spacing: float, lambda_c: float, A: int) -> Tuple[np.ndarray, int]:
    """(cos k phi_g, sin k phi_g), k = 1..A_pad-1, float32 [G][stride] for the lag-form scan; a ULA
    with positions m*d has steering phase m*phi_g, phi_g = 2 pi d sin(az_g) / lambda."""
    ap = padded_antennas(A)
    stride = (2 * (ap - 1) + 3) & ~3
    az = np.radians(grid_deg)
    tab = np.zeros((len(grid_deg), stride), dtype=np.float64)
    for k in range(1, ap):
        ph = 2 * np.pi * (k * spacing) * np.sin(az) / lambda_c
        tab[:, 2 * (k - 1)] = np.cos(ph)
        tab[:, 2 * (k - 1) + 1] = np.sin(ph)
    return tab.astype(np.float32), stride


def f16_split(x32: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """x (float32, |x| <= 1) -> (hi, lo) fp16 with x = hi + lo up to 2^-22 |x| (+ 3e-8 where lo is subnormal)."""
    x32 = np.asarray(x32, dtype=np.float32)
    hi = x32.astype(np.float16)
    lo = (x32 - hi.astype(np.float32)).astype(np.float16)
    return hi, lo


def scan_mma_table(scan: np.ndarray, G: int, A: int) -> Tuple[np.ndarray, int]:
    """B-operand fragments of the tensor-core angle scan (mma.sync m16n8k16 fp16, operands split hi + lo and packed along
    K as [hi(8) | lo(8)] per k-step) for a grid that is symmetric about 0: grid pairs (g, G-1-g), g < ceil(G/2), share
    cos(k phi_g) and differ in the sign of sin(k phi_g).
    Layout uint32 [ntiles][matrix: cos, sin][kstep][lane 32][2] = (packed b_hi, packed b_lo): a word holds the table values
    of lags 8 s + 2 (lane % 4) + 1 (low half) and + 2 (high half) for pair 8 j + lane / 4 (0 beyond the last lag / pair).
    Returned as a float32 view of those words (the C ABI passes it as const float*).  `scan` is the float32
    [G][stride] table of scan_table()."""
    ap = padded_antennas(A)
    assert ap in (8, 16)
    npairs = (G + 1) // 2
    nt = (npairs + 7) // 8
    K, KS = ap, ap // 8
    T = np.zeros((2, K, nt * 8), dtype=np.float32)
    for k in range(K):
        lag = k + 1
        if lag <= ap - 1:
            T[0, k, :npairs] = scan[:npairs, 2 * (lag - 1)]
            T[1, k, :npairs] = scan[:npairs, 2 * (lag - 1) + 1]
    hi, lo = f16_split(T)
    hi16, lo16 = hi.view(np.uint16).astype(np.uint32), lo.view(np.uint16).astype(np.uint32)
    out = np.zeros((nt, 2, KS, 32, 2), dtype=np.uint32)
    lane = np.arange(32)
    for j in range(nt):
        n = 8 * j + lane // 4
        for s_ in range(KS):
            k0 = 8 * s_ + 2 * (lane % 4)
            out[j, :, s_, :, 0] = hi16[:, k0, n] | (hi16[:, k0 + 1, n] << 16)
            out[j, :, s_, :, 1] = lo16[:, k0, n] | (lo16[:, k0 + 1, n] << 16)
    return out.reshape(-1).view(np.float32), nt


def scan_tc_table(scan: np.ndarray, G: int, A: int, pairs_per_half: int = 32) -> Tuple[np.ndarray, int]:
    """B operands of the tcgen05 angle scan (csrc/rs_angles.cu, angles_tc5_kernel) for a grid symmetric about 0.
    Grid pairs (g, G-1-g) are cut into jobs of 32 (the UMMA N).  Per job: the cos table then the sin table, each as KC
    chunks of [32 pairs x 16 K-slots] fp16 in the canonical K-major no-swizzle UMMA layout -- element (n, k) at byte
    (n // 8) * 256 + (k // 8) * 128 + (n % 8) * 16 + (k % 8) * 2.  K packing of the hi / lo split (x = hi + lo, both fp16):
        A <= 8 : chunk 0 = [hi(8) | hi(8)] (against A = [a_hi | a_lo]), chunk 1 = [lo(8) | 0] (against [a_hi | 0])
        A <= 16: chunk 0 = hi(16) (against a_hi), chunk 1 = hi(16) (against a_lo), chunk 2 = lo(16) (against a_hi)
    K-slot j holds lag j + 1; the last slot multiplies a constant one in the A operand: 0 for the grid pairs and -16384 in
    the cos table for the padding columns of the last job, so that the scan needs no mask there.
    Returns (uint8 bytes, number of halves)."""
    ap = padded_antennas(A)
    assert ap in (8, 16)
    npairs = (G + 1) // 2
    nh = (npairs + pairs_per_half - 1) // pairs_per_half
    kc = 2 if ap == 8 else 3
    T = np.zeros((2, ap, nh * pairs_per_half), dtype=np.float32)         # [cos / sin][slot][pair]
    for k in range(ap - 1):
        T[0, k, :npairs] = scan[:npairs, 2 * k]
        T[1, k, :npairs] = scan[:npairs, 2 * k + 1]
    T[0, ap - 1, npairs:] = -16384.0
    hi, lo = f16_split(T)
    out = np.zeros((nh, 2, kc, pairs_per_half * 16), dtype=np.float16)
    n = np.arange(pairs_per_half)
    for h in range(nh):
        cols = h * pairs_per_half + n
        for part in range(2):
            if ap == 8:
                chunks = [np.concatenate([hi[part, :, cols].T, hi[part, :, cols].T], axis=0),       # [16 slots][48]
                          np.concatenate([lo[part, :, cols].T, np.zeros((8, pairs_per_half), np.float16)], axis=0)]
            else:
                chunks = [hi[part, :, cols].T, hi[part, :, cols].T, lo[part, :, cols].T]
            for c, ch in enumerate(chunks):
                for k in range(16):
                    off = (n // 8) * 128 + (k // 8) * 64 + (n % 8) * 8 + (k % 8)                      # in fp16 elements
                    out[h, part, c, off] = ch[k]
    return out.reshape(-1).view(np.uint8), nh


STEER_TC_POINTS = 96          # grid points per N-half of the steering GEMM (192 accumulator columns: Re, Im interleaved)


def steer_tc_table(steer: np.ndarray) -> Tuple[np.ndarray, int]:
    """B operands of the tcgen05 steering GEMM (csrc/rs_angles.cu, music_tc_kernel; A > 16).
    y_g = a_g^H s is a real contraction over K = 2 A: the A operand row of a cell is, per chunk of 8 antennas, the 16 K slots
    [Re s (8) | Im s (8)]; column 2 j of an N-half is Re y_g, column 2 j + 1 is Im y_g of grid point g = 96 h + j:
        Re y_g = sum_m  Re a_m Re s_m + Im a_m Im s_m        Im y_g = sum_m  Re a_m Im s_m - Im a_m Re s_m.
    Layout: [halves][chunks][hi, lo][192 columns x 16 slots] fp16, every [192 x 16] block in the canonical K-major
    no-swizzle UMMA layout (element (n, k) at (n // 8) * 256 + (k // 8) * 128 + (n % 8) * 16 + (k % 8) * 2 bytes);
    x = hi + lo as in scan_tc_table.  steer is the c128 [A][G] table of `steering`; grid points and antennas beyond the
    table are zero columns / slots.  Returns (uint8 bytes, number of halves)."""
    A, G = steer.shape
    nh = (G + STEER_TC_POINTS - 1) // STEER_TC_POINTS
    nc = (A + 7) // 8
    st = np.zeros((nc * 8, nh * STEER_TC_POINTS), dtype=np.complex128)
    st[:A, :G] = steer
    N = 2 * STEER_TC_POINTS
    T = np.zeros((nh, nc, N, 16), dtype=np.float32)                          # [half][chunk][column][slot]
    for h in range(nh):
        blk = st[:, h * STEER_TC_POINTS:(h + 1) * STEER_TC_POINTS].reshape(nc, 8, STEER_TC_POINTS)   # [chunk][m][j]
        re, im = blk.real.transpose(0, 2, 1), blk.imag.transpose(0, 2, 1)    # [chunk][j][m]
        T[h, :, 0::2, 0:8] = re
        T[h, :, 0::2, 8:16] = im
        T[h, :, 1::2, 0:8] = -im
        T[h, :, 1::2, 8:16] = re
    hi, lo = f16_split(T)
    n = np.arange(N)[:, None]
    k = np.arange(16)[None, :]
    off = ((n // 8) * 128 + (k // 8) * 64 + (n % 8) * 8 + (k % 8)).reshape(-1)   # in fp16 elements
    out = np.zeros((nh, nc, 2, N * 16), dtype=np.float16)
    out[:, :, 0, off] = hi.reshape(nh, nc, -1)
    out[:, :, 1, off] = lo.reshape(nh, nc, -1)
    return out.reshape(-1).view(np.uint8), nh


def grid_cos_sin(grid_deg: np.ndarray) -> np.ndarray:
    """(cos, sin) of np.radians(grid) -- what velocity_solver.py:94-97 evaluates per target. f64 [G][2]."""
    az = np.radians(grid_deg)
    return np.stack([np.cos(az), np.sin(az)], axis=1)


def power_threshold(threshold_db: float) -> float:
    """p > thr  <=>  10 log10(p + 1e-12) > threshold_db   (dechirp.py:238, 252)."""
    return float(10.0 ** (threshold_db / 10.0) - 1e-12)


def range_axis(range_resolution: float, range_bins: int) -> np.ndarray:
    return np.linspace(0, range_resolution * range_bins, range_bins)          # dechirp.py:241


def doppler_axis(sampling_rate: float, doppler_bins: int) -> np.ndarray:
    return np.linspace(-sampling_rate / 2, sampling_rate / 2, doppler_bins)   # dechirp.py:242


def range_gate(range_resolution: float, range_bins: int, min_range: float, max_range: float) -> np.ndarray:
    r = range_axis(range_resolution, range_bins)
    return ((r >= min_range) & (r <= max_range)).astype(np.uint8)              # dechirp.py:263
