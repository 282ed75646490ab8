"""Empty stand-in for matplotlib.pyplot (tests only)."""
