"""Empty stand-in so the reference's unmodified scripts import where matplotlib is not installed (tests only)."""
