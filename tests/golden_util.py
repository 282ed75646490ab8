"""Helpers shared by the CPU (oracle) and GPU (CUDA path) golden tests."""
import ast
import os

import numpy as np

from oracle import radar_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SCENE = np.array([(8.0, np.radians(0.0), -10.0, 0.0), (12.0, np.radians(30.0), -8.0, 0.0),
                  (16.0, np.radians(-20.0), -6.0, 0.0), (20.0, np.radians(10.0), -3.0, 0.0),
                  (25.0, np.radians(-40.0), 0.0, 0.0)])

CASE_NAMES = ["c1_default", "ref_default_400x64", "c4_sparse", "small_hamming_nodc", "a16_blackman",
              "lownoise_400x32"]


def load_case(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"), allow_pickle=False))
    cfg = ast.literal_eval(str(g["meta"]))
    return g, cfg


def params_of(cfg) -> orc.RadarParams:
    return orc.RadarParams(chirp_duration=cfg["S"] / 10e6, num_chirps=cfg["C"], num_antennas=cfg["A"],
                           window_type=cfg["window"], dc_removal=cfg["dc"], noise_power=cfg["noise"])


def make_input(cfg) -> np.ndarray:
    """The complex64 cube the golden outputs were computed from (see oracle/make_golden.py)."""
    np.random.seed(cfg["seed"])
    return orc.synthesize_frame(params_of(cfg), SCENE).astype(np.complex64)


def check_input(g, cube64):
    chk = np.array([cube64.real.astype(np.float64).sum(), cube64.imag.astype(np.float64).sum(),
                    np.abs(cube64.astype(np.complex128)).sum()])
    np.testing.assert_allclose(chk, g["cube_checksum"], rtol=1e-12)
