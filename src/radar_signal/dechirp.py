"""Drop-in for radar-slam's src/radar_signal/dechirp.py -- same import path, B200 (sm_100a) implementation.
No __init__.py on purpose: the reference tree has none, so `src` is a PEP 420 namespace package and putting
this repo first on PYTHONPATH makes the reference's unmodified scripts pick this module."""
from radar_slam_b200.compat.dechirp import SignalPreprocessor, process_frame, main, logger  # noqa: F401

if __name__ == "__main__":
    main()
