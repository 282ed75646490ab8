"""Host-side constant tables of the tensor-core scans (radar_slam_b200/tables.py) against their definitions, on the CPU:
the UMMA operand layouts are decoded again and the contraction they encode is compared with the fp64 expression."""
import numpy as np

from radar_slam_b200 import tables


def _umma_block(flat16, rows):
    """[rows x 16] fp16 block in the canonical K-major no-swizzle layout -> array [rows][16]."""
    n = np.arange(rows)[:, None]
    k = np.arange(16)[None, :]
    off = (n // 8) * 128 + (k // 8) * 64 + (n % 8) * 8 + (k % 8)
    return flat16[off]


def test_steering_gemm_table_encodes_the_steering_product():
    rng = np.random.default_rng(3)
    for A, res in ((192, 1.0), (40, 0.5), (17, 2.0)):
        grid = tables.azimuth_grid((-90, 90), res)
        lam = 3e8 / 77e9
        steer = tables.steering(grid, np.arange(A) * 0.5 * lam, lam)
        raw, nh = tables.steer_tc_table(steer)
        nc = (A + 7) // 8
        assert nh == (len(grid) + 95) // 96 and raw.size == nh * nc * 2 * 192 * 16 * 2
        v = raw.view(np.float16).reshape(nh, nc, 2, 192 * 16)
        s = rng.standard_normal(A) + 1j * rng.standard_normal(A)
        sp = np.zeros(nc * 8, dtype=np.complex128)
        sp[:A] = s
        y = np.zeros((nh, 192))
        for h in range(nh):
            for c in range(nc):
                b = _umma_block(v[h, c, 0].astype(np.float64), 192) + _umma_block(v[h, c, 1].astype(np.float64), 192)
                a_row = np.concatenate([sp[8 * c:8 * c + 8].real, sp[8 * c:8 * c + 8].imag])     # [Re s (8) | Im s (8)]
                y[h] += b @ a_row
        got = (y[:, 0::2] + 1j * y[:, 1::2]).reshape(-1)
        want = steer.conj().T @ s                                                                # a_g^H s
        assert np.abs(got[:len(grid)] - want).max() <= 2e-6 * np.abs(want).max()                # hi + lo: ~2^-22 per entry
        assert np.all(got[len(grid):] == 0)                                                      # padding columns


def test_scan_table_pads_the_last_job_with_the_constant_slot():
    for A, res in ((8, 1.0), (16, 1.0), (6, 2.0)):
        grid = tables.azimuth_grid((-90, 90), res)
        lam = 3e8 / 77e9
        scan, stride = tables.scan_table(grid, 0.5 * lam, lam, A)
        raw, nh = tables.scan_tc_table(scan, len(grid), A)
        ap = tables.padded_antennas(A)
        kc = 2 if ap == 8 else 3
        npairs = (len(grid) + 1) // 2
        v = raw.view(np.float16).reshape(nh, 2, kc, 32 * 16)
        for h in range(nh):
            cos_hi = _umma_block(v[h, 0, 0].astype(np.float64), 32)           # chunk 0 of the cos table: hi parts
            for n in range(32):
                pair = 32 * h + n
                slot = cos_hi[n, ap - 1]                                      # the K slot that multiplies the constant one
                assert slot == (0.0 if pair < npairs else -16384.0)
                if pair < npairs and A > 1:
                    assert abs(cos_hi[n, 0] - scan[pair, 0]) <= 1e-3          # lag 1: cos(phi_pair), fp16 hi part
