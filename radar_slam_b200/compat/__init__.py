"""Drop-in implementations of the reference's hot-path classes, running on the
CUDA library.  The modules under ``src/`` at the repo root re-export these under the reference's import
paths so scripts/run_ego_motion_pipeline.py and the reference's tests import them unmodified."""
