"""radar_slam_b200 -- B200-native (sm_100a) implementation of radar-slam's per-frame hot path.

    from radar_slam_b200 import RadarConfig, FramePipeline

The drop-in replacements of the reference's modules live under ``src/`` at the repo root (same
import paths as the reference) and are thin re-exports of ``radar_slam_b200.compat``.
"""
from .pipeline import RadarConfig, FramePipeline, Detections   # noqa: F401
from ._lib import RadarSlamError                               # noqa: F401

__version__ = "0.1.0"
