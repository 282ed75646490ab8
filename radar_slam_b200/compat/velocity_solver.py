"""Drop-in for the reference's src/velocity_solver/velocity_solver.py (VelocitySolver,
estimate_velocity_from_angles, CLI) -- velocity_solver.py:20-467.

The reference minimises the phase cost with scipy differential_evolution (seed 42) twice.  The cost
is linear in the six motion parameters (velocity_solver.py:84-113, 171-174), i.e. a convex quadratic,
so the optimiser's fixed point is the bounded least-squares solution over the reference's box
(+-50, +-50, +-10 m/s and +-10 rad/s, :216, :250-251).  That solution is computed on the GPU in fp64
(rs_velocity_ls6: normal equations + active-set bounded solve).  Parameters whose design column is
identically zero -- v_z and omega for the targets solve_velocity builds, because elevation is 0 and
position = range * direction (:334-342) -- are unobservable; the reference returns RNG-dependent values
for them, this implementation returns 0.
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _device
from .lazy import LazyRecords, column_of, records_of

logger = logging.getLogger(__name__)


def _as_list(target_info):
    """targets as a record sequence (LazyRecords, the reference's object array of dicts, or a list)."""
    return records_of(target_info)


class VelocitySolver:
    def __init__(self, fc: float = 77e9, lambda_c: float = None, num_antennas: int = 8, antenna_spacing: float = None,
                 optimization_method: str = 'differential_evolution', max_iterations: int = 1000,
                 tolerance: float = 1e-6):
        self.fc = fc
        self.c = 3e8
        self.lambda_c = lambda_c or (self.c / self.fc)
        self.num_antennas = num_antennas
        self.antenna_spacing = antenna_spacing or (self.lambda_c / 2)
        self.optimization_method = optimization_method
        self.max_iterations = max_iterations
        self.tolerance = tolerance
        self.antenna_positions = np.arange(self.num_antennas) * self.antenna_spacing
        logger.info("Initialized velocity solver:")
        logger.info(f"  Wavelength: {self.lambda_c*1000:.2f} mm")
        logger.info(f"  Optimization method: {optimization_method}")
        if optimization_method != 'differential_evolution':
            # velocity_solver.py:255-262: any other method is an UNBOUNDED scipy.optimize.minimize from x0 = 0 (or the
            # initial guess) with maxiter / tol; here every method gets the box-bounded least-squares minimiser DE converges to
            logger.warning(f"optimization_method={optimization_method!r}: this path always returns the box-bounded least-squares "
                           "solution (the reference's non-DE branch is an unbounded local minimisation; max_iterations, "
                           "tolerance and initial_guess do not apply)")

    # ---- model helpers (vectorised restatements with the reference's signatures)
    def compute_phase_difference_model(self, target_positions: np.ndarray, target_angles: np.ndarray,
                                       velocity: np.ndarray, angular_velocity: np.ndarray, dt: float) -> np.ndarray:
        target_positions = np.asarray(target_positions, dtype=float).reshape(-1, 3)
        target_angles = np.asarray(target_angles, dtype=float).reshape(-1, 2)
        az, el = target_angles[:, 0], target_angles[:, 1]
        direction = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], axis=1)
        rel = np.asarray(velocity, dtype=float)[None, :] + np.cross(
            np.broadcast_to(np.asarray(angular_velocity, dtype=float), target_positions.shape), target_positions)
        return (4 * np.pi * np.sum(rel * direction, axis=1) * dt) / self.lambda_c

    def compute_observed_phase_differences(self, rds_data: np.ndarray, target_info: List[Dict]) -> np.ndarray:
        targets = _as_list(target_info)
        if isinstance(targets, LazyRecords) and len(targets):
            sig = np.asarray(targets.column('spatial_signature'))[:, :2].astype(complex)
        else:
            sig = np.array([[t['spatial_signature'][0], t['spatial_signature'][1]] for t in targets],
                           dtype=complex).reshape(-1, 2)
        return np.angle(sig[:, 1] * np.conj(sig[:, 0]))

    def cost_function(self, motion_params: np.ndarray, target_positions: np.ndarray, target_angles: np.ndarray,
                      observed_phases: np.ndarray, dt: float) -> float:
        motion_params = np.asarray(motion_params, dtype=float)
        predicted = self.compute_phase_difference_model(target_positions, target_angles, motion_params[:3],
                                                        motion_params[3:], dt)
        return float(np.sum((np.asarray(observed_phases) - predicted) ** 2))

    # ---- solver
    def _solve_box(self, pos, ang, y, dt, nvar):
        from .. import _lib
        pipe = _device.pipeline(fc=self.fc)
        dev = pipe.device
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)   # noqa: E731
        dpos, dang, dy = t(pos), t(ang), t(y)
        lo = t(np.array([-50.0, -50.0, -10.0, -10.0, -10.0, -10.0]))
        hi = t(np.array([50.0, 50.0, 10.0, 10.0, 10.0, 10.0]))
        out = torch.empty(7, dtype=torch.float64, device=dev)
        pred = torch.empty(len(y), dtype=torch.float64, device=dev)
        k = 4 * np.pi * dt / self.lambda_c
        pipe._call("rs_velocity_ls6", dpos.data_ptr(), dang.data_ptr(), dy.data_ptr(), len(y), float(k),
                   lo.data_ptr(), hi.data_ptr(), nvar, out.data_ptr(), pred.data_ptr(), pipe.stream)
        o = out.cpu().numpy()
        return o[:6].copy(), float(o[6]), pred.cpu().numpy()

    def two_step_optimization(self, target_positions: np.ndarray, target_angles: np.ndarray,
                              observed_phases: np.ndarray, dt: float,
                              initial_guess: Optional[np.ndarray] = None) -> Dict:
        """velocity_solver.py:178-307 with both optimiser calls replaced by the bounded least-squares solve
        they converge to (13 result keys, same failure value for N < 3)."""
        from scipy.optimize import OptimizeResult
        N = len(target_positions)
        if N < 3:
            logger.warning("Insufficient targets for optimization")
            return {'success': False, 'message': 'Insufficient targets'}
        pos = np.asarray(target_positions, dtype=float).reshape(N, 3)
        ang = np.asarray(target_angles, dtype=float).reshape(N, 2)
        y = np.asarray(observed_phases, dtype=float).reshape(N)
        logger.info("Step 1: Solving for translational velocity...")
        x1, f1, _ = self._solve_box(pos, ang, y, dt, 3)
        result_trans = OptimizeResult(x=x1[:3].copy(), fun=f1, success=True, nit=0, nfev=0,
                                      message='bounded least squares (closed form)')
        logger.info(f"Step 1 result: v_trans = {x1[:3]}")
        logger.info("Step 2: Refining with full 6-DoF motion...")
        x2, f2, predicted_phases = self._solve_box(pos, ang, y, dt, 6)
        result_full = OptimizeResult(x=x2.copy(), fun=f2, success=True, nit=0, nfev=0,
                                     message='bounded least squares (closed form)')
        velocity_est, angular_velocity_est = x2[:3].copy(), x2[3:].copy()
        residuals = y - predicted_phases
        rmse = np.sqrt(np.mean(residuals ** 2))
        max_residual = np.max(np.abs(residuals))
        results = {
            'success': True, 'velocity': velocity_est, 'angular_velocity': angular_velocity_est, 'cost': f2,
            'rmse': rmse, 'max_residual': max_residual, 'residuals': residuals, 'predicted_phases': predicted_phases,
            'observed_phases': y, 'num_targets': N, 'step1_result': result_trans, 'step2_result': result_full,
        }
        logger.info("Optimization complete:")
        logger.info(f"  Velocity: {velocity_est}")
        logger.info(f"  Angular velocity: {angular_velocity_est}")
        logger.info(f"  RMSE: {rmse:.6f}")
        logger.info(f"  Max residual: {max_residual:.6f}")
        return results

    def solve_velocity(self, rds_data: np.ndarray, target_info: List[Dict], dt: float = 0.1,
                       initial_guess: Optional[np.ndarray] = None) -> Dict:
        """velocity_solver.py:309-355."""
        targets = _as_list(target_info)
        range_m = column_of(targets, 'range_m', float).reshape(-1)
        az = column_of(targets, 'azimuth_rad', float).reshape(-1)
        el = np.zeros_like(az)                       # "Assume ground level" (velocity_solver.py:334)
        target_positions = np.stack([range_m * np.cos(el) * np.cos(az), range_m * np.cos(el) * np.sin(az),
                                     range_m * np.sin(el)], axis=1).reshape(-1, 3)
        target_angles = np.stack([az, el], axis=1).reshape(-1, 2)
        observed_phases = self.compute_observed_phase_differences(rds_data, targets) if len(targets) else np.zeros(0)
        return self.two_step_optimization(target_positions, target_angles, observed_phases, dt, initial_guess)

    def visualize_results(self, results: Dict, save_path: Optional[str] = None) -> None:
        if not results['success']:
            logger.warning("Cannot visualize failed optimization")
            return
        import matplotlib.pyplot as plt
        fig, axes = plt.subplots(2, 2, figsize=(12, 10))
        axes[0, 0].bar(['vx', 'vy', 'vz'], results['velocity'])
        axes[0, 0].set_title('Translational Velocity')
        axes[0, 1].bar(['wx', 'wy', 'wz'], results['angular_velocity'])
        axes[0, 1].set_title('Rotational Velocity')
        axes[1, 0].plot(results['residuals'], 'o-', alpha=0.7)
        axes[1, 0].set_title('Phase Residuals')
        axes[1, 1].scatter(results['observed_phases'], results['predicted_phases'], alpha=0.7)
        axes[1, 1].set_title('Predicted vs Observed')
        plt.tight_layout()
        if save_path:
            plt.savefig(save_path, dpi=150, bbox_inches='tight')
        plt.show()


def estimate_velocity_from_angles(angles_path: str, rds_path: str, output_path: str, radar_params: Dict = None,
                                  dt: float = 0.1) -> Dict:
    """velocity_solver.py:418-467."""
    angles_data = np.load(angles_path, allow_pickle=True)
    target_info = _as_list(angles_data['targets'])
    rds = np.load(rds_path)
    logger.info(f"Loaded {len(target_info)} targets")
    logger.info(f"RDS shape: {rds.shape}")
    if radar_params is None:
        radar_params = {'fc': 77e9, 'lambda_c': 3e8 / 77e9, 'num_antennas': 8}
    solver = VelocitySolver(**radar_params)
    results = solver.solve_velocity(rds, target_info, dt)
    np.savez(output_path, **results)
    logger.info(f"Velocity estimation complete: {results['success']}")
    if results['success']:
        logger.info(f"  Velocity: {results['velocity']}")
        logger.info(f"  Angular velocity: {results['angular_velocity']}")
        logger.info(f"  RMSE: {results['rmse']:.6f}")
    return results


def main(argv=None):
    import argparse
    parser = argparse.ArgumentParser(description='Estimate velocity from angles')
    parser.add_argument('--angles', required=True, help='Path to angles file')
    parser.add_argument('--rds', required=True, help='Path to RDS file')
    parser.add_argument('--out', required=True, help='Output path for velocity')
    parser.add_argument('--dt', type=float, default=0.1, help='Time step (s)')
    args = parser.parse_args(argv)
    results = estimate_velocity_from_angles(args.angles, args.rds, args.out, dt=args.dt)
    print(f"Velocity estimation complete: {results}")
