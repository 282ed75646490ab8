"""Synthetic FMCW cubes for benchmarks and smoke runs.

Same signal model as the reference simulator (scripts/simulate_raw.py:102-221): the scatterer term
is chirp-invariant (simulate_raw.py:190-209 never uses the chirp start time), so it is one
[A, S] plane per frame; rs_synthesize_frames evaluates it in fp64 on the device and streams plane + Philox noise
(simulate_raw.py:216-219) into the cube.  Frame k depends only on (seed, k), so every rank of a sharded run can
generate exactly its own frames.  scatterer_term() is the host fp64 version of the same plane (tests, CPU arms).
"""
from __future__ import annotations

import numpy as np
import torch

from .pipeline import RadarConfig
from . import tables

DEFAULT_SCENE = np.array([(8.0, np.radians(0.0), -10.0, 0.0), (12.0, np.radians(30.0), -8.0, 0.0),
                          (16.0, np.radians(-20.0), -6.0, 0.0), (20.0, np.radians(10.0), -3.0, 0.0),
                          (25.0, np.radians(-40.0), 0.0, 0.0)])


def scatterer_term(cfg: RadarConfig, scatterers: np.ndarray = DEFAULT_SCENE) -> np.ndarray:
    """[A, S] complex128: sum over scatterers of amplitude * steering * (delayed chirp * conj(ref))."""
    S = cfg.samples_per_chirp
    t = np.linspace(0, cfg.chirp_duration, S)
    ref = tables.reference_chirp(cfg.fc, cfg.chirp_rate, cfg.chirp_duration, S)
    pos = np.arange(cfg.num_antennas) * cfg.spacing
    out = np.zeros((cfg.num_antennas, S), dtype=np.complex128)
    for rng_m, az, rcs_db, vr in np.atleast_2d(scatterers)[:, :4]:
        if rng_m <= 0 or not np.isfinite([rng_m, az, rcs_db, vr]).all():
            continue
        amp = np.sqrt(10 ** (rcs_db / 10)) / (4 * np.pi * rng_m ** 2)
        steer = amp * np.exp(1j * (4 * np.pi * vr * cfg.fc / tables.C0 + 2 * np.pi * pos * np.sin(az) / cfg.lambda_c))
        td = t - 2 * rng_m / tables.C0
        valid = (td >= 0) & (td <= cfg.chirp_duration)
        if valid.any():
            tv = td[valid]
            delayed = np.exp(1j * (2 * np.pi * (cfg.fc * tv + 0.5 * cfg.chirp_rate * tv ** 2)))
            out[:, valid] += steer[:, None] * (delayed * np.conj(ref[valid]))[None, :]
    return out


def synthesize_frames(cfg: RadarConfig, scatterers, frames: int, seed: int, noise_power: float = 0.01,
                      device=None, first_frame: int = 0, out: torch.Tensor = None) -> torch.Tensor:
    """FMCWRadarSimulator.synthesize_frame for `frames` frames on the device (rs_synthesize_frames).
    scatterers: [n, 4] (the same scene in every frame) or [frames, n, 4] = range_m, azimuth_rad, rcs_db, radial velocity.
    Returns complex64 [frames, A, C, S].  Frame k's noise depends only on (seed, first_frame + k)."""
    from . import _lib
    if not torch.cuda.is_available():
        raise _lib.RadarSlamError("radar_slam_b200 needs a CUDA device (no CPU fallback)")
    lib = _lib.load()
    device = torch.device(device or f"cuda:{torch.cuda.current_device()}")
    A, C, S = cfg.num_antennas, cfg.num_chirps, cfg.samples_per_chirp
    sc = np.asarray(scatterers, dtype=np.float64)
    if sc.ndim == 2:
        sc = np.broadcast_to(sc[None, :, :4], (frames,) + sc[:, :4].shape)
    sc = np.array(sc[:, :, :4], dtype=np.float64, order="C", copy=True)
    assert sc.shape[0] == frames
    n_max = sc.shape[1]
    with torch.cuda.device(device):
        sc_dev = torch.from_numpy(sc).to(device) if n_max else torch.zeros(1, dtype=torch.float64, device=device)
        pos = torch.from_numpy(np.arange(A) * cfg.spacing).to(device)
        plane = torch.empty((frames, A, S), dtype=torch.complex64, device=device)
        cube = out if out is not None else torch.empty((frames, A, C, S), dtype=torch.complex64, device=device)
        assert cube.is_contiguous() and cube.dtype == torch.complex64 and tuple(cube.shape) == (frames, A, C, S)
        _lib.check(lib.rs_synthesize_frames(sc_dev.data_ptr(), 0, n_max, cfg.fc, cfg.chirp_rate, cfg.chirp_duration,
                                            cfg.lambda_c, pos.data_ptr(), float(noise_power), int(seed) & (2 ** 64 - 1),
                                            int(first_frame), plane.data_ptr(), cube.data_ptr(), frames, A, C, S,
                                            torch.cuda.current_stream(device).cuda_stream), "rs_synthesize_frames")
    return cube


def synth_cubes(cfg: RadarConfig, frames: int, seed: int, noise_power: float = 0.01,
                scatterers: np.ndarray = DEFAULT_SCENE, device=None, first_frame: int = 0) -> torch.Tensor:
    """complex64 [frames, A, C, S] on the device: the benchmark workload (same scene in every frame, fresh noise per
    frame).  Any rank can produce frames [first_frame, first_frame + frames) of the same sequence."""
    return synthesize_frames(cfg, scatterers, frames, seed, noise_power, device, first_frame)
