"""Drop-in for radar-slam's src/algorithms/velocity_solver_improved.py -- same import path, CUDA path underneath
(SURVEY.md 8f3).  No __init__.py on purpose (PEP 420 namespace shadowing)."""
from radar_slam_b200.compat.velocity_solver_improved import (  # noqa: F401
    ImprovedVelocitySolver, estimate_velocity_improved, logger)
