"""Drop-in for radar-slam's src/angle_estimation/angle_estimation.py -- B200 (sm_100a) implementation."""
from radar_slam_b200.compat.angle_estimation import AngleEstimator, extract_angles_from_rds, main, logger  # noqa: F401

if __name__ == "__main__":
    main()
