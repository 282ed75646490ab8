"""Drop-in for radar-slam's src/robust_angle_estimation.py -- B200 (sm_100a) implementation."""
from radar_slam_b200.compat.robust_angle_estimation import RobustAngleEstimator, extract_angles_robust, main, logger  # noqa: F401

if __name__ == "__main__":
    main()
