"""Drop-in for radar-slam's src/algorithms/robust_angle_estimation.py (a byte-identical copy of
src/robust_angle_estimation.py in the reference; its tests import this path) -- B200 implementation."""
from radar_slam_b200.compat.robust_angle_estimation import RobustAngleEstimator, extract_angles_robust, main, logger  # noqa: F401

if __name__ == "__main__":
    main()
