"""Shared device plumbing of the legacy adapters: pipeline cache, fp64 upload / detect / snapshot helpers."""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import replace
from typing import Optional, Tuple

import numpy as np
import torch

from .. import _lib
from ..pipeline import FramePipeline, RadarConfig

_pipes: "OrderedDict[tuple, FramePipeline]" = OrderedDict()


def pipeline(**kw) -> FramePipeline:
    """A FramePipeline for these radar parameters (cached; parameters are read at call time, so mutating
    an attribute of a legacy object after construction behaves like the reference)."""
    key = tuple(sorted((k, (tuple(v) if isinstance(v, (list, tuple)) else v)) for k, v in kw.items()))
    p = _pipes.get(key)
    if p is None:
        p = FramePipeline(RadarConfig(**kw))
        _pipes[key] = p
        while len(_pipes) > 16:
            _pipes.popitem(last=False)
    return p


def rds_to_device(rds: np.ndarray, pipe: FramePipeline) -> torch.Tensor:
    """Reference-layout RDS [A, R, D] (host, any complex dtype) -> complex128 [A, R, D] on the device.

    Always a fresh upload of exactly the array the caller passed: the legacy API works on whatever RDS it is
    handed (edited in place, np.load-ed from a stage file, ...), one frame per call, so there is nothing to cache."""
    rds = np.asarray(rds)
    if rds.ndim != 3:
        raise ValueError("rds must be [num_antennas, range_bins, doppler_bins]")
    return torch.from_numpy(np.ascontiguousarray(rds, dtype=np.complex128)).to(pipe.device)


def detect_f64(pipe: FramePipeline, rds128: torch.Tensor, gate: np.ndarray, threshold_db: float):
    """extract_range_doppler_peaks (dechirp.py:235-263) in fp64 on the device.
    -> (antenna, range_bin, doppler_bin int64 arrays in the reference's order, power_db float64 [A, R, D])."""
    A, R, D = rds128.shape
    dev = pipe.device
    power_db = torch.empty((A, R, D), dtype=torch.float64, device=dev)
    row_count = torch.empty((A * R,), dtype=torch.int32, device=dev)
    row_off = torch.empty((A * R,), dtype=torch.int64, device=dev)
    total = torch.zeros((1,), dtype=torch.int64, device=dev)
    gate_dev = torch.from_numpy(np.ascontiguousarray(gate, dtype=np.uint8)).to(dev)
    cap = max(1024, (A * R * D) // 8)
    while True:
        keys = torch.empty((cap,), dtype=torch.int32, device=dev)
        pipe._call("rs_detect_f64", rds128.data_ptr(), gate_dev.data_ptr(), float(threshold_db), power_db.data_ptr(),
                   row_count.data_ptr(), row_off.data_ptr(), keys.data_ptr(), cap, total.data_ptr(), A, R, D, pipe.stream)
        n = int(total.item())
        if n <= cap:
            break
        cap = n                                   # plateau-heavy input: every cell can be a peak
    k = keys[:n].cpu().numpy().view(np.uint32).astype(np.int64)
    return k >> 24, (k >> 12) & 0xFFF, k & 0xFFF, power_db.cpu().numpy()


def keys_tensor(antenna, range_bin, doppler_bin, device) -> torch.Tensor:
    a = np.asarray(antenna, dtype=np.int64)
    r = np.asarray(range_bin, dtype=np.int64)
    d = np.asarray(doppler_bin, dtype=np.int64)
    key = ((a & 0xFF) << 24) | (r << 12) | d
    return torch.from_numpy(key.astype(np.uint32).view(np.int32)).to(device)


def signatures(pipe: FramePipeline, rds128: torch.Tensor, range_bin, doppler_bin) -> torch.Tensor:
    """Unit-energy snapshots complex128 [n, A] of cells of a complex128 RDS [A, R, D] (angle_estimation.py:83-88)."""
    A, R, D = rds128.shape
    n = len(range_bin)
    rb = torch.from_numpy(np.ascontiguousarray(range_bin, dtype=np.int32)).to(pipe.device)
    db = torch.from_numpy(np.ascontiguousarray(doppler_bin, dtype=np.int32)).to(pipe.device)
    out = torch.empty((n, A), dtype=torch.complex128, device=pipe.device)
    pipe._call("rs_signatures_c128", rds128.data_ptr(), rb.data_ptr(), db.data_ptr(), n, out.data_ptr(), A, R, D,
               pipe.stream)
    return out


def spectra(pipe: FramePipeline, sig128: torch.Tensor, steer128: torch.Tensor, method: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """fp64 pseudo-spectra [n, G] and first-index argmax [n] for snapshots complex128 [n, A]."""
    n, A = sig128.shape
    G = steer128.shape[1]
    out = torch.empty((n, G), dtype=torch.float64, device=pipe.device)
    aidx = torch.empty((n,), dtype=torch.int32, device=pipe.device)
    pipe._call("rs_spectra_f64", sig128.data_ptr(), steer128.data_ptr(), _lib.METHODS[method], n, A, G,
               out.data_ptr(), aidx.data_ptr(), pipe.stream)
    return out, aidx


def esprit(pipe: FramePipeline, sig128: torch.Tensor, scale: float) -> torch.Tensor:
    n, A = sig128.shape
    out = torch.empty((n,), dtype=torch.float64, device=pipe.device)
    pipe._call("rs_esprit_f64", sig128.data_ptr(), n, A, float(scale), out.data_ptr(), pipe.stream)
    return out
