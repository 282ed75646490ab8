"""Drop-in for radar-slam's scripts/simulate_raw.py (import name `simulate_raw`): with <repo>/scripts ahead of the
reference's scripts directory on PYTHONPATH, run_ego_motion_pipeline.py's `from simulate_raw import
FMCWRadarSimulator` (run_ego_motion_pipeline.py:32) picks the CUDA-path simulator."""
from radar_slam_b200.compat.simulate_raw import FMCWRadarSimulator, logger  # noqa: F401
