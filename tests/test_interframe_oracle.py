"""oracle/interframe_oracle.py against tests/golden/interframe_de.npz, which oracle/make_interframe_golden.py wrote by
running the REFERENCE's own ImprovedVelocitySolver (association, temporal phase differences, phase model, cost and
differential_evolution(seed=42)) on seeded synthetic target sets."""
import os

import numpy as np
import pytest

from oracle import interframe_oracle as ifo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "interframe_de.npz")
CASES = ["slow", "fast", "dense"]
DT, LAM = 0.1, 3e8 / 77e9


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.mark.parametrize("name", CASES)
def test_association_and_phases_match_the_reference(gold, name):
    g = gold
    cur = ifo.target_xy(g[f"{name}_cur_range"], g[f"{name}_cur_az"])
    prev = ifo.target_xy(g[f"{name}_prev_range"], g[f"{name}_prev_az"])
    match, dist = ifo.associate(cur, prev, 5.0)
    hit = np.nonzero(match >= 0)[0]
    assert np.array_equal(hit, g[f"{name}_match_cur"]) and np.array_equal(match[hit], g[f"{name}_match"])
    np.testing.assert_allclose(dist[hit], g[f"{name}_dist"], rtol=0, atol=1e-12)
    assert (match < 0).sum() == len(cur) - len(g[f"{name}_match"])     # clutter without a partner stays unmatched
    y = ifo.temporal_phase(g[f"{name}_cur_sig"][hit, 0], g[f"{name}_prev_sig"][match[hit], 0])
    np.testing.assert_allclose(y, g[f"{name}_y"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", CASES)
def test_model_and_cost_match_the_reference(gold, name):
    g = gold
    hit = g[f"{name}_match_cur"]
    rng, az = g[f"{name}_cur_range"][hit], g[f"{name}_cur_az"][hit]
    pos = np.stack([rng * np.cos(az), rng * np.sin(az), np.zeros_like(az)], axis=1)
    ang = np.stack([az, np.zeros_like(az)], axis=1)
    probes = g[f"{name}_probes"]
    np.testing.assert_allclose(ifo.phase_model(pos, ang, probes[0, :3], probes[0, 3:], DT, LAM), g[f"{name}_model"],
                               rtol=1e-12, atol=1e-9)
    got = np.array([ifo.wrapped_cost(m, pos, ang, g[f"{name}_y"], DT, LAM) for m in probes])
    np.testing.assert_allclose(got, g[f"{name}_probe_cost"], rtol=1e-10, atol=1e-9)
    # the last probe is where the reference's differential evolution stopped
    assert abs(got[-1] - float(g[f"{name}_de_cost"])) < 1e-6


def test_the_references_optimiser_does_not_find_the_true_motion(gold):
    """Why the CUDA path replaces differential evolution by a global lattice search: on a noise-level problem whose true
    motion has a cost of ~0.01, the reference's answer is a local minimum tens of m/s away with a cost of ~10."""
    g = gold
    hit = g["slow_match_cur"]
    rng, az = g["slow_cur_range"][hit], g["slow_cur_az"][hit]
    pos = np.stack([rng * np.cos(az), rng * np.sin(az), np.zeros_like(az)], axis=1)
    ang = np.stack([az, np.zeros_like(az)], axis=1)
    true = np.concatenate([g["slow_v_true"], np.zeros(4)])
    assert ifo.wrapped_cost(true, pos, ang, g["slow_y"], DT, LAM) < 0.1 < 5.0 < float(g["slow_de_cost"])
    assert np.abs(g["slow_de_velocity"][:2] - g["slow_v_true"]).max() > 10.0
