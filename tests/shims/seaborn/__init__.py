"""Empty stand-in for seaborn (tests only)."""
