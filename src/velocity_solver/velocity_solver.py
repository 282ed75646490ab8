"""Drop-in for radar-slam's src/velocity_solver/velocity_solver.py -- B200 (sm_100a) implementation."""
from radar_slam_b200.compat.velocity_solver import VelocitySolver, estimate_velocity_from_angles, main, logger  # noqa: F401

if __name__ == "__main__":
    main()
