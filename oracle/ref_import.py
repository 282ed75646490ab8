"""TEST INFRASTRUCTURE ONLY -- import the *real* reference classes when /root/reference exists.

The reference is pure Python but imports matplotlib and h5py at module top
(dechirp.py:14, angle_estimation.py:15, robust_angle_estimation.py:12,
velocity_solver.py:12, simulate_raw.py:15); neither is installed in this image, so
empty stand-in modules are injected first.  Only plotting and HDF5 reading are lost.

/root/reference does not exist on the GPU box: callers must check ``available()``
and skip.  Nothing in the product imports this module.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("RADAR_SLAM_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "radar_signal", "dechirp.py"))


def _stub(name: str) -> None:
    if name in sys.modules:
        return
    try:
        if importlib.util.find_spec(name) is not None:
            return
    except (ImportError, ValueError):
        pass
    sys.modules[name] = types.ModuleType(name)


def _load(path: str, modname: str):
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


_cache: dict = {}


def load():
    """Return a namespace with the reference classes, loaded by file path under private
    module names so they never collide with this repo's own drop-in ``src`` modules."""
    if _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError(f"reference tree not present at {REF_ROOT}")
    for n in ("matplotlib", "matplotlib.pyplot", "h5py", "seaborn"):
        _stub(n)
    if isinstance(sys.modules.get("matplotlib"), types.ModuleType) and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    j = os.path.join
    ns = types.SimpleNamespace()
    ns.dechirp = _load(j(REF_ROOT, "src/radar_signal/dechirp.py"), "_rsref_dechirp")
    ns.angle = _load(j(REF_ROOT, "src/angle_estimation/angle_estimation.py"), "_rsref_angle")
    ns.robust = _load(j(REF_ROOT, "src/robust_angle_estimation.py"), "_rsref_robust")
    ns.velocity = _load(j(REF_ROOT, "src/velocity_solver/velocity_solver.py"), "_rsref_velocity")
    ns.simulate = _load(j(REF_ROOT, "scripts/simulate_raw.py"), "_rsref_simulate")
    ns.SignalPreprocessor = ns.dechirp.SignalPreprocessor
    ns.AngleEstimator = ns.angle.AngleEstimator
    ns.RobustAngleEstimator = ns.robust.RobustAngleEstimator
    ns.VelocitySolver = ns.velocity.VelocitySolver
    ns.FMCWRadarSimulator = ns.simulate.FMCWRadarSimulator
    import logging
    for m in ("_rsref_dechirp", "_rsref_angle", "_rsref_robust", "_rsref_velocity", "_rsref_simulate"):
        logging.getLogger(m).setLevel(logging.ERROR)
    logging.getLogger().setLevel(logging.WARNING)
    _cache["ns"] = ns
    return ns
