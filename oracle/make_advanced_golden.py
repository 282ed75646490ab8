"""TEST INFRASTRUCTURE ONLY -- writes tests/golden/advanced_de.npz by running the REFERENCE's own
AdvancedVelocityOptimizer (src/algorithms/advanced_velocity_optimization.py: regularised cost :153-223, initial guesses
:260-341, differential_evolution(seed=42) per run :343-408, multi-run driver and adaptive bounds :410-525, :94-151)
on seeded synthetic association lists.  Needs /root/reference; run once in the build container:

    python -m oracle.make_advanced_golden
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402


def associations(rs, n, v_true, dt, lam, noise, jitter):
    """Associated target pairs of a static scene seen from a platform moving with v_true (planar)."""
    rng = rs.uniform(5.0, 60.0, n)
    az = rs.uniform(-1.2, 1.2, n)
    k = 4 * np.pi * dt / lam
    y = k * (v_true[0] * np.cos(az) + v_true[1] * np.sin(az)) + noise * rs.randn(n)
    y = np.arctan2(np.sin(y), np.cos(y))
    out = []
    for i in range(n):
        cur = {'range_m': float(rng[i]), 'azimuth_rad': float(az[i])}
        prev = {'range_m': float(rng[i] + jitter * rs.randn()), 'azimuth_rad': float(az[i] + 0.01 * jitter * rs.randn())}
        out.append({'current': cur, 'previous': prev, 'distance': 0.0, 'temporal_phase_diff': float(y[i])})
    return out


def main():
    ref_import.load()
    mod = ref_import._load(os.path.join(ref_import.REF_ROOT, "src/algorithms/advanced_velocity_optimization.py"),
                           "_rsref_advanced")
    import logging
    logging.getLogger("_rsref_advanced").setLevel(logging.ERROR)
    dt, lam = 0.1, 3e8 / 77e9
    out = {}
    cases = [("slow", 12, (0.004, -0.003), 0.03, 0.3, 21, None),
             ("prev", 10, (3.2, -1.1), 0.05, 0.3, 22, np.array([3.0, -1.0, 0.2, 0.01, -0.02, 0.3]))]
    for name, n, v_true, noise, jitter, seed, previous in cases:
        rs = np.random.RandomState(seed)
        assoc = associations(rs, n, v_true, dt, lam, noise, jitter)
        opt = mod.AdvancedVelocityOptimizer(fc=77e9, lambda_c=lam, num_antennas=4, num_optimization_runs=3, use_parallel=False)
        pos = np.array([[a['current']['range_m'] * np.cos(a['current']['azimuth_rad']),
                         a['current']['range_m'] * np.sin(a['current']['azimuth_rad']), 0.0] for a in assoc])
        ang = np.array([[a['current']['azimuth_rad'], 0.0] for a in assoc])
        y = np.array([a['temporal_phase_diff'] for a in assoc])
        # probes through every branch of the regulariser: inside, |v| > 0.8 max, |w| > 0.8 max, both large, v_z
        probes = np.concatenate([
            rs.uniform(-1, 1, (6, 6)) * np.array([50, 50, 10, 10, 10, 10]),
            np.array([[45.0, 10.0, 1.0, 0.1, 0.2, 0.3], [1.0, 2.0, 0.5, 6.0, 5.0, 4.0], [25.0, 5.0, -2.0, 4.0, 3.0, 2.0],
                      [0.004, -0.003, 0.0, 0.0, 0.0, 0.0], [3.2, -1.1, 0.0, 0.0, 0.0, 0.0]])])
        probe_cost = np.array([opt.compute_regularized_cost_function(m, pos, ang, y, dt, None) for m in probes])
        pm = np.array([3.0, -1.0, 0.2, 0.01, -0.02, 0.3])
        probe_cost_prev = np.array([opt.compute_regularized_cost_function(m, pos, ang, y, dt, pm) for m in probes])
        model = opt._compute_phase_difference_model(pos, ang, probes[0, :3], probes[0, 3:], dt)
        np.random.seed(seed)
        guesses = np.array(opt.generate_multiple_initial_guesses(assoc, dt))
        np.random.seed(seed)
        res = opt.run_robust_optimization(assoc, dt, previous)
        print(name, "success", res.get('success'), "v", res.get('velocity'), "w", res.get('angular_velocity'), "cost",
              res.get('cost'), "true", v_true, "runs", res.get('successful_runs'), flush=True)
        # the reference reports failure when differential_evolution stops at maxiter without meeting tol (:379-397: every
        # run then returns cost inf and run_robust_optimization says 'All optimization runs failed'); the point DE reached
        # is recorded by replaying the reference's exact call (:372-377) on a fresh optimiser's bounds
        from scipy.optimize import differential_evolution
        fresh = mod.AdvancedVelocityOptimizer(fc=77e9, lambda_c=lam, num_antennas=4, num_optimization_runs=3, use_parallel=False)
        bounds = fresh.adaptive_bounds['velocity_bounds'] + fresh.adaptive_bounds['angular_velocity_bounds']
        de = differential_evolution(lambda m: fresh.compute_regularized_cost_function(m, pos, ang, y, dt, previous), bounds,
                                    maxiter=1000, tol=1e-6, seed=42, workers=1)
        print(name, "replayed DE: success", de.success, "nit", de.nit, "x", de.x, "fun", de.fun, flush=True)
        out.update({f"{name}_replay_x": de.x, f"{name}_replay_cost": np.array(float(de.fun)),
                    f"{name}_replay_success": np.array(bool(de.success)), f"{name}_replay_nit": np.array(de.nit)})
        pack = lambda key, which: np.array([a[which][key] for a in assoc])                          # noqa: E731
        out.update({
            f"{name}_cur_range": pack('range_m', 'current'), f"{name}_cur_az": pack('azimuth_rad', 'current'),
            f"{name}_prev_range": pack('range_m', 'previous'), f"{name}_prev_az": pack('azimuth_rad', 'previous'),
            f"{name}_y": y, f"{name}_v_true": np.array(v_true), f"{name}_previous": previous if previous is not None else np.zeros(0),
            f"{name}_probes": probes, f"{name}_probe_cost": probe_cost, f"{name}_probe_cost_prev": probe_cost_prev,
            f"{name}_probe_prev_motion": pm, f"{name}_model": model, f"{name}_guesses": guesses, f"{name}_seed": np.array(seed),
            f"{name}_ref_success": np.array(bool(res.get('success'))),
        })
        if res.get('success'):
            out.update({
                f"{name}_de_velocity": res['velocity'], f"{name}_de_angular": res['angular_velocity'],
                f"{name}_de_cost": np.array(float(res['cost'])), f"{name}_de_rmse": np.array(float(res['rmse'])),
                f"{name}_de_runs": np.array([res['num_optimization_runs'], res['successful_runs']]),
                f"{name}_bounds_v": np.array(opt.adaptive_bounds['velocity_bounds'], dtype=float),
                f"{name}_bounds_acc": np.array(opt.adaptive_bounds['acceleration_bounds'], dtype=float),
            })
        # adaptive bounds after two updates (the second one has a history to difference)
        opt2 = mod.AdvancedVelocityOptimizer(fc=77e9, lambda_c=lam, num_antennas=4)
        opt2.update_adaptive_bounds(np.array([4.0, -2.0, 0.1]), np.array([0.0, 0.1, 0.2]), dt)
        opt2.update_adaptive_bounds(np.array([4.5, -1.0, 0.0]), np.array([0.05, 0.1, 0.1]), dt)
        for key, val in opt2.adaptive_bounds.items():
            out[f"{name}_upd_{key}"] = np.array(val, dtype=float)
    path = os.path.join(ROOT, "tests", "golden", "advanced_de.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
