"""Inter-frame velocity path on the GPU: rs_associate_targets / rs_wrapped_cost against the oracle,
the lattice search against the reference's own differential-evolution answers (tests/golden/interframe_de.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import interframe_oracle as ifo

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "interframe_de.npz")
DT, LAM = 0.1, 3e8 / 77e9


def _targets(g, name, which):
    rng, az, sig = g[f"{name}_{which}_range"], g[f"{name}_{which}_az"], g[f"{name}_{which}_sig"]
    return [{"range_m": float(rng[i]), "azimuth_rad": float(az[i]), "spatial_signature": sig[i]} for i in range(len(rng))]


@pytest.mark.parametrize("nc,npv,thr,seed", [(1, 1, 5.0, 0), (37, 52, 5.0, 1), (300, 280, 2.0, 2), (64, 64, 1e9, 3), (20, 30, 1e-9, 4)])
def test_association_matches_oracle(nc, npv, thr, seed):
    from radar_slam_b200 import _lib
    lib = _lib.load()
    rs = np.random.RandomState(seed)
    cur = rs.uniform(-40, 40, (nc, 2))
    prev = np.concatenate([cur[rs.permutation(nc)][: min(nc, npv)] + 0.5 * rs.randn(min(nc, npv), 2),
                           rs.uniform(-40, 40, (max(0, npv - nc), 2))])[:npv]
    if nc > 3 and npv > 3:
        prev[2] = prev[1]                                      # exact duplicates: equal distances, first index wins
    want_m, want_d = ifo.associate(cur, prev, thr)
    dev = "cuda"
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)       # noqa: E731
    c, p = t(cur, np.float64), t(prev, np.float64)
    ncd, npd = t(np.array([nc]), np.int32), t(np.array([npv]), np.int32)
    m = torch.empty(nc, dtype=torch.int32, device=dev)
    d = torch.empty(nc, dtype=torch.float64, device=dev)
    _lib.check(lib.rs_associate_targets(c.data_ptr(), ncd.data_ptr(), p.data_ptr(), npd.data_ptr(), thr, m.data_ptr(),
                                        d.data_ptr(), 1, nc, npv, torch.cuda.current_stream().cuda_stream))
    assert np.array_equal(m.cpu().numpy(), want_m)
    assert np.array_equal(d.cpu().numpy(), want_d)             # same fp64 expression: bit equal


@pytest.mark.parametrize("name", ["slow", "fast", "dense"])
def test_solver_class_against_the_references_run(name):
    from src.algorithms.velocity_solver_improved import ImprovedVelocitySolver
    g = np.load(GOLD)
    cur, prev = _targets(g, name, "cur"), _targets(g, name, "prev")
    solver = ImprovedVelocitySolver(fc=77e9, lambda_c=LAM, num_antennas=4)
    assoc = solver.associate_targets_across_frames(cur, prev)
    assert [a["current"] is cur[i] for a, i in zip(assoc, g[f"{name}_match_cur"])] == [True] * len(assoc)
    assert [a["previous"] is prev[j] for a, j in zip(assoc, g[f"{name}_match"])] == [True] * len(assoc)
    np.testing.assert_allclose([a["distance"] for a in assoc], g[f"{name}_dist"], rtol=0, atol=1e-12)
    y = solver.compute_observed_phase_differences(assoc)
    np.testing.assert_allclose(y, g[f"{name}_y"], rtol=0, atol=1e-12)

    hit = g[f"{name}_match_cur"]
    rng, az = g[f"{name}_cur_range"][hit], g[f"{name}_cur_az"][hit]
    pos = np.stack([rng * np.cos(az), rng * np.sin(az), np.zeros_like(az)], axis=1)
    ang = np.stack([az, np.zeros_like(az)], axis=1)
    probes = g[f"{name}_probes"]
    got = np.array([solver.cost_function(m, pos, ang, y, DT) for m in probes])          # rs_wrapped_cost
    np.testing.assert_allclose(got, g[f"{name}_probe_cost"], rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(solver.compute_phase_difference_model(pos, ang, probes[0, :3], probes[0, 3:], DT),
                               g[f"{name}_model"], rtol=1e-12, atol=1e-9)

    res = solver.solve_velocity_with_association(cur, prev, dt=DT)
    assert res["success"] and res["num_associations"] == len(assoc)
    assert set(res) == {"success", "velocity", "angular_velocity", "cost", "rmse", "max_residual", "residuals",
                        "predicted_phases", "observed_phases", "num_associations", "step1_result", "step2_result"}
    # never worse than what the reference's differential evolution returned, and it is the reference's own cost
    assert res["cost"] <= float(g[f"{name}_de_cost"]) + 1e-9
    full = np.concatenate([res["velocity"], res["angular_velocity"]])
    assert abs(ifo.wrapped_cost(full, pos, ang, y, DT, LAM) - res["cost"]) < 1e-9
    # ... and it is the true motion (all three scenes are static targets seen from a moving platform)
    assert np.abs(res["velocity"][:2] - g[f"{name}_v_true"]).max() < 2e-3
    assert res["cost"] < 0.02 * len(assoc) and np.all(res["angular_velocity"] == 0) and res["velocity"][2] == 0


def test_too_few_associations_and_empty_inputs():
    from src.algorithms.velocity_solver_improved import ImprovedVelocitySolver
    s = ImprovedVelocitySolver()
    t = [{"range_m": 10.0, "azimuth_rad": 0.1, "spatial_signature": np.ones(2, complex)}]
    assert s.associate_targets_across_frames(t, []) == [] and s.associate_targets_across_frames([], t) == []
    assert s.solve_velocity_with_association(t, []) == {"success": False, "message": "No target associations"}
    assoc = s.associate_targets_across_frames(t, t)
    assert len(assoc) == 1 and assoc[0]["distance"] == 0.0 and assoc[0]["temporal_phase_diff"] == 0.0
    assert s.two_step_optimization(assoc, 0.1) == {"success": False, "message": "Insufficient target associations"}
    assert np.array_equal(s.get_smart_initial_guess([], 0.1), np.zeros(6))
