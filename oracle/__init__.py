"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the radar-slam per-frame hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker or the reported CPU baseline.
The product path (``radar_slam_b200`` and the drop-in ``src/`` modules) never
imports this package and fails loudly when the CUDA library is missing.

Parity status: the reference's own tests pin no numeric result on this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself: ``oracle/make_golden.py`` imports the real classes from /root/reference
(under a matplotlib/h5py stub) and writes ``tests/golden/*.npz``; the restatement in
``oracle/radar_oracle.py`` is checked against those fixtures and, when
/root/reference is present, against the live reference classes.  The inter-frame solver
(``oracle/interframe_oracle.py``) is pinned the same way through ``tests/golden/interframe_de.npz``, written by
``oracle/make_interframe_golden.py`` from a run of the reference's own ``ImprovedVelocitySolver``.
"""
