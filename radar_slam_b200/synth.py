"""Synthetic FMCW cubes for benchmarks and smoke runs.

Same signal model as the reference simulator (scripts/simulate_raw.py:102-221): the scatterer term
is chirp-invariant (simulate_raw.py:190-209 never uses the chirp start time), so it is evaluated
once on the host in fp64 -- the chirp phase needs fp64 -- and broadcast; the complex Gaussian noise
(simulate_raw.py:216-219) is drawn on the device.  Frame k depends only on (seed, k-block), so
every rank of a sharded run can generate exactly its own frames.
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


def synth_cubes(cfg: RadarConfig, frames: int, seed: int, noise_power: float = 0.01,
                scatterers: np.ndarray = DEFAULT_SCENE, device=None, first_frame: int = 0,
                block: int = 64) -> torch.Tensor:
    """complex64 [frames, A, C, S] on the device.  Noise is generated in blocks of `block` frames with
    generator seed (seed, block index), so any rank can produce frames [first_frame, first_frame+frames)."""
    device = torch.device(device or f"cuda:{torch.cuda.current_device()}")
    A, C, S = cfg.num_antennas, cfg.num_chirps, cfg.samples_per_chirp
    cube = torch.empty((frames, A, C, S), dtype=torch.complex64, device=device)
    sig = torch.from_numpy(scatterer_term(cfg, scatterers).astype(np.complex64)).to(device)
    sigma = float(np.sqrt(noise_power))
    gen = torch.Generator(device=device)
    f = first_frame
    while f < first_frame + frames:
        b = f // block
        lo, hi = b * block, (b + 1) * block
        gen.manual_seed((seed << 20) + b)
        blk = torch.empty((block, A, C, S, 2), dtype=torch.float32, device=device)
        blk.normal_(0.0, sigma, generator=gen)
        s0, s1 = max(lo, first_frame), min(hi, first_frame + frames)
        cube[s0 - first_frame:s1 - first_frame] = torch.view_as_complex(blk[s0 - lo:s1 - lo])
        f = hi
    cube += sig[None, :, None, :]
    return cube
