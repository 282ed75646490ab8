"""AdvancedVelocityOptimizer on the GPU (advanced_velocity_optimization.py): the regularised cost
(rs_regularized_cost) against the reference's own values, host bookkeeping (adaptive bounds, initial guesses) against the
reference's, and the global search against the point the reference's differential evolution reached
(tests/golden/advanced_de.npz, written by oracle/make_advanced_golden.py from the reference's class)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "advanced_de.npz")
DT, LAM = 0.1, 3e8 / 77e9


def _assoc(g, name):
    out = []
    for i in range(len(g[f"{name}_y"])):
        out.append({"current": {"range_m": float(g[f"{name}_cur_range"][i]), "azimuth_rad": float(g[f"{name}_cur_az"][i])},
                    "previous": {"range_m": float(g[f"{name}_prev_range"][i]), "azimuth_rad": float(g[f"{name}_prev_az"][i])},
                    "distance": 0.0, "temporal_phase_diff": float(g[f"{name}_y"][i])})
    return out


@pytest.mark.parametrize("name", ["slow", "prev"])
def test_advanced_optimizer_against_the_references_run(name):
    from src.algorithms.advanced_velocity_optimization import AdvancedVelocityOptimizer, optimize_velocity_advanced
    g = np.load(GOLD)
    assoc = _assoc(g, name)
    opt = AdvancedVelocityOptimizer(fc=77e9, lambda_c=LAM, num_antennas=4, num_optimization_runs=3, use_parallel=False)
    rng, az = g[f"{name}_cur_range"], g[f"{name}_cur_az"]
    pos = np.stack([rng * np.cos(az), rng * np.sin(az), np.zeros_like(az)], axis=1)
    ang = np.stack([az, np.zeros_like(az)], axis=1)
    y = g[f"{name}_y"]
    probes = g[f"{name}_probes"]
    # the cost, through every branch of the regulariser, with and without a previous motion
    got = np.array([opt.compute_regularized_cost_function(m, pos, ang, y, DT) for m in probes])
    np.testing.assert_allclose(got, g[f"{name}_probe_cost"], rtol=1e-10, atol=1e-9)
    pm = g[f"{name}_probe_prev_motion"]
    got = np.array([opt.compute_regularized_cost_function(m, pos, ang, y, DT, pm) for m in probes])
    np.testing.assert_allclose(got, g[f"{name}_probe_cost_prev"], rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(opt._compute_phase_difference_model(pos, ang, probes[0, :3], probes[0, 3:], DT),
                               g[f"{name}_model"], rtol=1e-12, atol=1e-9)
    # host bookkeeping: same numpy RNG calls, same smart guess, same adaptive bounds
    np.random.seed(int(g[f"{name}_seed"]))
    np.testing.assert_allclose(np.array(opt.generate_multiple_initial_guesses(assoc, DT)), g[f"{name}_guesses"], rtol=0, atol=1e-12)
    o2 = AdvancedVelocityOptimizer(fc=77e9, lambda_c=LAM, num_antennas=4)
    o2.update_adaptive_bounds(np.array([4.0, -2.0, 0.1]), np.array([0.0, 0.1, 0.2]), DT)
    o2.update_adaptive_bounds(np.array([4.5, -1.0, 0.0]), np.array([0.05, 0.1, 0.1]), DT)
    for key, val in o2.adaptive_bounds.items():
        np.testing.assert_allclose(np.array(val, dtype=float), g[f"{name}_upd_{key}"], rtol=0, atol=1e-12)

    previous = g[f"{name}_previous"] if len(g[f"{name}_previous"]) else None
    res = opt.run_robust_optimization(assoc, DT, previous)
    assert res["success"] and res["num_associations"] == len(assoc)
    assert set(res) == {"success", "velocity", "angular_velocity", "cost", "rmse", "max_residual", "residuals",
                        "predicted_phases", "observed_phases", "num_associations", "num_optimization_runs", "successful_runs",
                        "best_initial_guess", "all_results"}
    assert res["num_optimization_runs"] == 3 and res["successful_runs"] == 3 and len(res["all_results"]) == 3
    assert set(res["all_results"][0]) == {"success", "motion_params", "cost", "iterations", "initial_guess"}
    # never worse than the point the reference's differential evolution reached, measured with the reference's own cost
    assert res["cost"] <= float(g[f"{name}_replay_cost"]) + 1e-9
    full = np.concatenate([res["velocity"], res["angular_velocity"]])
    fresh = AdvancedVelocityOptimizer(fc=77e9, lambda_c=LAM, num_antennas=4)
    assert abs(fresh.compute_regularized_cost_function(full, pos, ang, y, DT, previous) - res["cost"]) < 1e-9
    # ... and it is the true planar motion of the scene
    assert np.abs(res["velocity"][:2] - g[f"{name}_v_true"]).max() < 5e-3
    assert res["cost"] < 0.02 * len(assoc) + (0.0 if previous is None else 1e-3)
    # the adaptive bounds moved with the result (:497), like the reference's would
    assert len(opt.velocity_history) == 1 and np.array_equal(opt.velocity_history[0], res["velocity"])
    r2 = optimize_velocity_advanced(assoc, DT, {"fc": 77e9, "lambda_c": LAM, "num_antennas": 4}, previous)
    assert r2["success"] and np.allclose(r2["velocity"], res["velocity"])


def test_advanced_optimizer_too_few_associations():
    from src.algorithms.advanced_velocity_optimization import AdvancedVelocityOptimizer
    opt = AdvancedVelocityOptimizer()
    a = [{"current": {"range_m": 10.0, "azimuth_rad": 0.1}, "previous": {"range_m": 10.1, "azimuth_rad": 0.1},
          "temporal_phase_diff": 0.1}] * 2
    assert opt.run_robust_optimization(a, 0.1) == {"success": False, "message": "Insufficient target associations"}
    assert np.array_equal(opt._generate_smart_initial_guess([], 0.1), np.zeros(6))
    assert opt.adaptive_bounds["velocity_bounds"] == [(-50.0, 50.0)] * 3
