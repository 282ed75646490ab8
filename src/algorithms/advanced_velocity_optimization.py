"""Drop-in for radar-slam's src/algorithms/advanced_velocity_optimization.py -- B200 (sm_100a) implementation."""
from radar_slam_b200.compat.advanced_velocity_optimization import (AdvancedVelocityOptimizer, optimize_velocity_advanced,  # noqa: F401
                                                                   main, logger)

if __name__ == "__main__":
    main()
