"""Drop-in for radar-slam's src/algorithms/advanced_velocity_optimization.py on the CUDA path.

Same class, constructor and method signatures, attributes and result keys as the reference (advanced_velocity_optimization.py
:24-525).  The regularised cost (:153-223) runs on the device for batches of candidate motions (rs_regularized_cost); the
adaptive bounds (:94-151) and the initial-guess generators (:260-341, same numpy RNG calls, so the global random state
advances as in the reference) are host bookkeeping.

The optimiser is NOT differential evolution.  The reference's run_single_optimization (:343-408) calls
differential_evolution(seed=42) on the six adaptive bounds and ignores the initial guess it is handed, so its "multiple
runs" all return the same point; and because the cost wraps its residual while the model's phase slope is 4 pi dt / lambda,
that point is whichever of ~1e7 local minima the population reaches.  Here one global search serves every run: with
elevation 0 and position = range * direction (:433-441) the wrapped part depends on (v_x, v_y) only, so the whole velocity
box is searched on a lattice finer than a basin (rs_wrapped_lattice_search), the best tiles are polished on the device
(rs_wrapped_gn_polish, with the temporal regulariser when previous_motion is given), v_z and the angular velocity take
the values that minimise the regularisers (0, or the shrunk previous motion), and the candidates are ranked by the
reference's own full cost.  The returned cost is <= the cost of the reference's answer
(tests/test_gpu_advanced_velocity.py checks it against a committed run of the reference's class).
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional

import numpy as np
import torch

from .. import _lib
from . import _device
from .velocity_solver_improved import wrapped_global_search

logger = logging.getLogger(__name__)


class AdvancedVelocityOptimizer:
    def __init__(self, fc: float = 77e9, lambda_c: float = None, num_antennas: int = 8, antenna_spacing: float = None,
                 max_velocity: float = 50.0, max_angular_velocity: float = 10.0, regularization_weight: float = 0.01,
                 num_optimization_runs: int = 3, use_parallel: bool = True):
        self.fc = fc
        self.c = 3e8
        self.lambda_c = lambda_c or (self.c / self.fc)
        self.num_antennas = num_antennas
        self.antenna_spacing = antenna_spacing or (self.lambda_c / 2)
        self.max_velocity = max_velocity
        self.max_angular_velocity = max_angular_velocity
        self.regularization_weight = regularization_weight
        self.num_optimization_runs = num_optimization_runs
        self.use_parallel = use_parallel
        self.antenna_positions = np.arange(self.num_antennas) * self.antenna_spacing
        self.velocity_history = []
        self.angular_velocity_history = []
        self.adaptive_bounds = self._initialize_adaptive_bounds()
        # not in the reference: the lattice search's knobs
        self.lattice_points_per_period = 6.0
        self.polish_candidates = 48
        logger.info("Initialized advanced velocity optimizer:")
        logger.info(f"  Max velocity: {max_velocity} m/s")
        logger.info(f"  Max angular velocity: {max_angular_velocity} rad/s")
        logger.info(f"  Regularization weight: {regularization_weight}")
        logger.info(f"  Optimization runs: {num_optimization_runs}")
        logger.info(f"  Parallel processing: {use_parallel}")

    # ---- adaptive bounds (host bookkeeping, advanced_velocity_optimization.py:85-151)
    def _initialize_adaptive_bounds(self) -> Dict:
        return {
            'velocity_bounds': [(-self.max_velocity, self.max_velocity)] * 3,
            'angular_velocity_bounds': [(-self.max_angular_velocity, self.max_angular_velocity)] * 3,
            'acceleration_bounds': [(-20, 20)] * 3,
            'angular_acceleration_bounds': [(-5, 5)] * 3,
        }

    def update_adaptive_bounds(self, current_velocity: np.ndarray, current_angular_velocity: np.ndarray,
                               dt: float = 0.1) -> None:
        self.velocity_history.append(current_velocity.copy())
        self.angular_velocity_history.append(current_angular_velocity.copy())
        max_history = 10
        if len(self.velocity_history) > max_history:
            self.velocity_history = self.velocity_history[-max_history:]
            self.angular_velocity_history = self.angular_velocity_history[-max_history:]
        if len(self.velocity_history) >= 2:
            vel_changes = np.diff(self.velocity_history, axis=0)
            ang_vel_changes = np.diff(self.angular_velocity_history, axis=0)
            max_acceleration = np.max(np.abs(vel_changes) / dt) if dt > 0 else 20.0
            max_angular_acceleration = np.max(np.abs(ang_vel_changes) / dt) if dt > 0 else 5.0
            safety_factor = 2.0
            self.adaptive_bounds['acceleration_bounds'] = [(-max_acceleration * safety_factor,
                                                            max_acceleration * safety_factor)] * 3
            self.adaptive_bounds['angular_acceleration_bounds'] = [(-max_angular_acceleration * safety_factor,
                                                                    max_angular_acceleration * safety_factor)] * 3
            current_speed = np.linalg.norm(current_velocity)
            if current_speed > 0:
                direction = current_velocity / current_speed
                velocity_expansion = min(10.0, current_speed * 0.5)
                for i in range(3):
                    if direction[i] > 0:
                        self.adaptive_bounds['velocity_bounds'][i] = (
                            -self.max_velocity, min(self.max_velocity, current_velocity[i] + velocity_expansion))
                    else:
                        self.adaptive_bounds['velocity_bounds'][i] = (
                            max(-self.max_velocity, current_velocity[i] - velocity_expansion), self.max_velocity)

    # ---- cost and model
    @staticmethod
    def _dev():
        return _device.pipeline(fc=77e9).device          # any pipeline: only the library handle and the device matter

    def _costs(self, motions: np.ndarray, pos: np.ndarray, ang: np.ndarray, y: np.ndarray, dt: float,
               previous_motion: Optional[np.ndarray]) -> np.ndarray:
        """compute_regularized_cost_function for a batch of motions [Q, 6] on the device."""
        dev = self._dev()
        lib = _lib.load()
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)      # noqa: E731
        m, p, a, yy = t(np.reshape(motions, (-1, 6))), t(np.reshape(pos, (-1, 3))), t(np.reshape(ang, (-1, 2))), t(np.reshape(y, -1))
        prev = t(np.reshape(previous_motion, 6)) if previous_motion is not None else None
        out = torch.empty(m.shape[0], dtype=torch.float64, device=dev)
        _lib.check(lib.rs_regularized_cost(m.data_ptr(), p.data_ptr(), a.data_ptr(), yy.data_ptr(), int(yy.numel()),
                                           int(m.shape[0]), 4 * np.pi * dt / self.lambda_c, float(self.max_velocity),
                                           float(self.max_angular_velocity), float(self.regularization_weight),
                                           _lib.ptr(prev), out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                   "rs_regularized_cost")
        return out.cpu().numpy()

    def compute_regularized_cost_function(self, motion_params: np.ndarray, target_positions: np.ndarray,
                                          target_angles: np.ndarray, observed_phases: np.ndarray, dt: float,
                                          previous_motion: Optional[np.ndarray] = None) -> float:
        return float(self._costs(np.asarray(motion_params, dtype=float), np.asarray(target_positions, dtype=float),
                                 np.asarray(target_angles, dtype=float), np.asarray(observed_phases, dtype=float), dt,
                                 previous_motion)[0])

    def _compute_phase_difference_model(self, target_positions: np.ndarray, target_angles: np.ndarray,
                                        velocity: np.ndarray, angular_velocity: np.ndarray, dt: float) -> np.ndarray:
        pos = np.asarray(target_positions, dtype=float).reshape(-1, 3)
        ang = np.asarray(target_angles, dtype=float).reshape(-1, 2)
        d = np.stack([np.cos(ang[:, 1]) * np.cos(ang[:, 0]), np.cos(ang[:, 1]) * np.sin(ang[:, 0]), np.sin(ang[:, 1])], axis=1)
        rel = np.asarray(velocity, dtype=float)[None, :] + np.cross(np.asarray(angular_velocity, dtype=float)[None, :], pos)
        return (4 * np.pi * np.sum(rel * d, axis=1) * dt) / self.lambda_c

    # ---- initial guesses (the same numpy RNG calls as the reference, :260-341)
    def generate_multiple_initial_guesses(self, target_associations: List[Dict], dt: float) -> List[np.ndarray]:
        initial_guesses = [self._generate_smart_initial_guess(target_associations, dt), np.zeros(6)]
        for _ in range(self.num_optimization_runs - 2):
            initial_guesses.append(np.array([
                np.random.uniform(-self.max_velocity * 0.5, self.max_velocity * 0.5),
                np.random.uniform(-self.max_velocity * 0.5, self.max_velocity * 0.5),
                np.random.uniform(-5, 5),
                np.random.uniform(-self.max_angular_velocity * 0.5, self.max_angular_velocity * 0.5),
                np.random.uniform(-self.max_angular_velocity * 0.5, self.max_angular_velocity * 0.5),
                np.random.uniform(-self.max_angular_velocity * 0.5, self.max_angular_velocity * 0.5)]))
        return initial_guesses

    def _generate_smart_initial_guess(self, target_associations: List[Dict], dt: float) -> np.ndarray:
        if not target_associations:
            return np.zeros(6)
        vel = []
        for a in target_associations:
            c, p = a['current'], a['previous']
            cp = np.array([c['range_m'] * np.cos(c['azimuth_rad']), c['range_m'] * np.sin(c['azimuth_rad']), 0])
            pp = np.array([p['range_m'] * np.cos(p['azimuth_rad']), p['range_m'] * np.sin(p['azimuth_rad']), 0])
            vel.append((cp - pp) / dt)
        med = np.median(np.array(vel), axis=0)
        return np.concatenate([np.append(-med[:2], 0), np.array([0, 0, 0])])

    # ---- the optimiser
    def _global_minimum(self, target_positions, target_angles, observed_phases, dt, previous_motion):
        """(motion [6], cost, lattice points) minimising the regularised cost over the adaptive bounds."""
        pos = np.asarray(target_positions, dtype=float).reshape(-1, 3)
        ang = np.asarray(target_angles, dtype=float).reshape(-1, 2)
        y = np.asarray(observed_phases, dtype=float).reshape(-1)
        k = 4 * np.pi * dt / self.lambda_c
        vb, wb = self.adaptive_bounds['velocity_bounds'], self.adaptive_bounds['angular_velocity_bounds']
        w = self.regularization_weight
        prev = None if previous_motion is None else np.asarray(previous_motion, dtype=float).reshape(6)
        planar = np.all(ang[:, 1] == 0) and np.all(pos[:, 2] == 0) and \
            np.allclose(pos[:, 0] * np.sin(ang[:, 0]), pos[:, 1] * np.cos(ang[:, 0]), rtol=0, atol=1e-9 * (1 + np.abs(pos).max()))
        if not planar:
            raise NotImplementedError("AdvancedVelocityOptimizer: the lattice search covers the reference's planar model "
                                      "(elevation 0, position = range * direction, advanced_velocity_optimization.py:433-441)")
        reg = 0.1 * w if prev is not None else 0.0
        centre = (prev[0], prev[1]) if prev is not None else (0.0, 0.0)
        cand, _, npts = wrapped_global_search(self._dev(), np.cos(ang[:, 0]), np.sin(ang[:, 0]), y, k, (vb[0], vb[1]),
                                              reg_lattice=0.0, reg_polish=reg, centre=centre,
                                              points_per_period=self.lattice_points_per_period,
                                              polish_candidates=self.polish_candidates)
        # v_z and w enter the regularisers only: 10 w v_z^2 + 0.1 w (v_z - p_z)^2 and 0.1 w |w - p_w|^2
        clip = lambda x, b: float(min(max(x, b[0]), b[1]))                                        # noqa: E731
        vz = clip(prev[2] * 0.1 / 10.1, vb[2]) if prev is not None else clip(0.0, vb[2])
        wv = [clip(prev[3 + i], wb[i]) if prev is not None else clip(0.0, wb[i]) for i in range(3)]
        full = np.concatenate([cand, np.tile(np.array([vz] + wv), (len(cand), 1))], axis=1)
        # a second family with zero rotation: the speed / rate product (term 4) can make the previous rotation a loss
        alt = full.copy()
        alt[:, 3:] = [clip(0.0, wb[i]) for i in range(3)]
        allc = np.concatenate([full, alt], axis=0)
        costs = self._costs(allc, pos, ang, y, dt, prev)
        b = int(np.argmin(costs))
        return allc[b], float(costs[b]), npts

    def run_single_optimization(self, initial_guess: np.ndarray, target_positions: np.ndarray, target_angles: np.ndarray,
                                observed_phases: np.ndarray, dt: float, previous_motion: Optional[np.ndarray] = None) -> Dict:
        try:
            motion, cost, npts = self._global_minimum(target_positions, target_angles, observed_phases, dt, previous_motion)
            return {'success': True, 'motion_params': motion, 'cost': cost, 'iterations': npts, 'initial_guess': initial_guess}
        except Exception as e:            # advanced_velocity_optimization.py:399-407
            return {'success': False, 'motion_params': initial_guess, 'cost': float('inf'), 'iterations': 0,
                    'initial_guess': initial_guess, 'error': str(e)}

    def run_robust_optimization(self, target_associations: List[Dict], dt: float,
                                previous_motion: Optional[np.ndarray] = None) -> Dict:
        if len(target_associations) < 3:
            logger.warning("Insufficient target associations for optimization")
            return {'success': False, 'message': 'Insufficient target associations'}
        rng = np.array([a['current']['range_m'] for a in target_associations], dtype=float)
        az = np.array([a['current']['azimuth_rad'] for a in target_associations], dtype=float)
        el = np.zeros_like(az)                                                   # "Assume ground level" (:433)
        target_positions = np.stack([rng * np.cos(el) * np.cos(az), rng * np.cos(el) * np.sin(az), rng * np.sin(el)], axis=1)
        target_angles = np.stack([az, el], axis=1)
        observed_phases = np.array([a['temporal_phase_diff'] for a in target_associations])
        initial_guesses = self.generate_multiple_initial_guesses(target_associations, dt)
        # the reference's runs differ only in an initial guess its optimiser never reads: one search serves them all
        first = self.run_single_optimization(initial_guesses[0], target_positions, target_angles, observed_phases, dt,
                                             previous_motion)
        results = [first] + [dict(first, initial_guess=g) if first['success'] else dict(first, motion_params=g, initial_guess=g)
                             for g in initial_guesses[1:]]
        successful_results = [r for r in results if r['success']]
        if not successful_results:
            logger.warning("All optimization runs failed")
            return {'success': False, 'message': 'All optimization runs failed'}
        best_result = min(successful_results, key=lambda x: x['cost'])
        motion_params = best_result['motion_params']
        velocity, angular_velocity = motion_params[:3], motion_params[3:]
        predicted_phases = self._compute_phase_difference_model(target_positions, target_angles, velocity, angular_velocity, dt)
        residuals = observed_phases - predicted_phases
        residuals = np.arctan2(np.sin(residuals), np.cos(residuals))
        self.update_adaptive_bounds(velocity, angular_velocity, dt)
        return {
            'success': True, 'velocity': velocity, 'angular_velocity': angular_velocity, 'cost': best_result['cost'],
            'rmse': np.sqrt(np.mean(residuals ** 2)), 'max_residual': np.max(np.abs(residuals)), 'residuals': residuals,
            'predicted_phases': predicted_phases, 'observed_phases': observed_phases,
            'num_associations': len(target_associations), 'num_optimization_runs': len(results),
            'successful_runs': len(successful_results), 'best_initial_guess': best_result['initial_guess'],
            'all_results': results,
        }


def optimize_velocity_advanced(target_associations: List[Dict], dt: float = 0.1, radar_params: Dict = None,
                               previous_motion: Optional[np.ndarray] = None) -> Dict:
    """advanced_velocity_optimization.py:527-566."""
    if radar_params is None:
        radar_params = {'fc': 77e9, 'lambda_c': 3e8 / 77e9, 'num_antennas': 8}
    optimizer = AdvancedVelocityOptimizer(**radar_params)
    results = optimizer.run_robust_optimization(target_associations, dt, previous_motion)
    logger.info(f"Advanced velocity optimization complete: {results['success']}")
    if results['success']:
        logger.info(f"  Velocity: {results['velocity']}")
        logger.info(f"  Angular velocity: {results['angular_velocity']}")
        logger.info(f"  RMSE: {results['rmse']:.6f}")
        logger.info(f"  Successful runs: {results['successful_runs']}/{results['num_optimization_runs']}")
    return results


def main(argv=None):
    import argparse
    parser = argparse.ArgumentParser(description='Advanced velocity optimization')
    parser.add_argument('--associations', required=True, help='Path to target associations file')
    parser.add_argument('--out', required=True, help='Output path for velocity')
    parser.add_argument('--dt', type=float, default=0.1, help='Time step (s)')
    args = parser.parse_args(argv)
    associations_data = np.load(args.associations, allow_pickle=True)
    results = optimize_velocity_advanced(list(associations_data['associations']), args.dt)
    print(f"Advanced velocity optimization complete: {results}")
