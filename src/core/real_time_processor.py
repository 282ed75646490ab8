"""Drop-in for radar-slam's src/core/real_time_processor.py -- same import path, the worker runs the CUDA path instead
of the reference's placeholder.  No __init__.py on purpose (PEP 420 namespace shadowing)."""
from radar_slam_b200.compat.real_time_processor import (  # noqa: F401
    ProcessingFrame, FrameBuffer, ParallelTargetProcessor, RealTimeProcessor, RealTimeVelocityEstimator,
    create_real_time_estimator, logger)
