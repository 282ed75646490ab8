"""Empty stand-in for h5py (tests only)."""
