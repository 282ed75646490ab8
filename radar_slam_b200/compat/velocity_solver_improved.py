"""Drop-in for radar-slam's src/algorithms/velocity_solver_improved.py on the CUDA path.

Same class, constructor and method signatures and result keys as the reference.  Association, the temporal phase
differences, the phase model and the cost are the reference's (association and the cost run on the device:
rs_associate_targets, rs_wrapped_cost).  The optimiser is NOT differential evolution: the reference's cost wraps its
residual to (-pi, pi] while the model's phase slope is 4 pi dt / lambda (322 rad per m/s at dt = 0.1 s, 77 GHz), so the
cost has a local minimum every ~2 cm/s per target direction over a +-50 m/s box, and `differential_evolution(seed=42)`
(velocity_solver_improved.py:389-420) returns whichever of them its population happens to reach (scipy-version
dependent).  two_step_optimization here searches the whole box on a lattice finer than a basin
(rs_wrapped_lattice_search, ~1e9 points on the GPU), polishes the best tiles in fp64 inside their basins
(rs_wrapped_gn_polish, Gauss-Newton on the device) and returns the global minimiser; its cost is <= the cost of the reference's answer (tests/test_gpu_interframe.py checks that against a
committed run of the reference's own class).  v_z and w do not enter the reference's planar model (elevation 0,
position = range * direction) and are returned as 0, which is where the regulariser puts them.
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional

import numpy as np
import torch

from .. import _lib
from . import _device
from .lazy import LazyRecords, column_of, records_of

logger = logging.getLogger(__name__)


def wrapped_global_search(dev, c: np.ndarray, s: np.ndarray, y: np.ndarray, k: float, bounds, reg_lattice: float,
                          reg_polish: float, centre, points_per_period: float = 6.0, polish_candidates: int = 48):
    """Global search of  sum_i wrap(y_i - k (v_x c_i + v_y s_i))^2 (+ reg |v - centre|^2)  over a (v_x, v_y) box, all on the
    device: rs_wrapped_lattice_search evaluates the cost on a lattice finer than a basin (step 2 pi / (k points_per_period))
    and keeps the best point of every tile; the `polish_candidates` best tiles are refined inside their basins by
    rs_wrapped_gn_polish (fp64 Gauss-Newton).  Returns (candidates [K, 2], their polished costs [K], lattice points)."""
    lib = _lib.load()
    (x_lo, x_hi), (y_lo, y_hi) = bounds
    x_lo, x_hi, y_lo, y_hi = float(x_lo), float(x_hi), float(y_lo), float(y_hi)
    h = 2 * np.pi / abs(k) / points_per_period
    nx, ny = int(np.floor((x_hi - x_lo) / h)) + 1, int(np.floor((y_hi - y_lo) / h)) + 1
    tx, ty = _lib.C.c_int(), _lib.C.c_int()
    _lib.check(lib.rs_wrapped_lattice_tiles(nx, ny, _lib.C.byref(tx), _lib.C.byref(ty)), "rs_wrapped_lattice_tiles")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)      # noqa: E731
    st = torch.cuda.current_stream(dev).cuda_stream
    ax, by, yc = t(k * c / (2 * np.pi)), t(k * s / (2 * np.pi)), t(y / (2 * np.pi))
    ntile = tx.value * ty.value
    tc = torch.empty(ntile, dtype=torch.float32, device=dev)
    tix = torch.empty(ntile, dtype=torch.int32, device=dev)
    tiy = torch.empty(ntile, dtype=torch.int32, device=dev)
    _lib.check(lib.rs_wrapped_lattice_search(ax.data_ptr(), by.data_ptr(), yc.data_ptr(), len(y), x_lo, y_lo, h, nx, ny,
                                             float(reg_lattice), tc.data_ptr(), tix.data_ptr(), tiy.data_ptr(), st),
               "rs_wrapped_lattice_search")
    kbest = min(polish_candidates, ntile)
    _, idx = torch.topk(tc, kbest, largest=False)
    v = torch.stack([x_lo + tix[idx].double() * h, y_lo + tiy[idx].double() * h], dim=1).contiguous()
    cost = torch.empty(kbest, dtype=torch.float64, device=dev)
    cd, sd, yd = t(c), t(s), t(y)
    _lib.check(lib.rs_wrapped_gn_polish(cd.data_ptr(), sd.data_ptr(), yd.data_ptr(), len(y), float(k), float(reg_polish),
                                        float(centre[0]), float(centre[1]), x_lo, x_hi, y_lo, y_hi, v.data_ptr(),
                                        cost.data_ptr(), kbest, 20, st), "rs_wrapped_gn_polish")
    return v.cpu().numpy(), cost.cpu().numpy(), int(nx) * int(ny)


class _Result(dict):
    """scipy.optimize.OptimizeResult look-alike (attribute access) for 'step1_result' / 'step2_result'."""
    __getattr__ = dict.get


class ImprovedVelocitySolver:
    def __init__(self, fc: float = 77e9, lambda_c: float = None, num_antennas: int = 8, antenna_spacing: float = None,
                 optimization_method: str = 'differential_evolution', max_iterations: int = 1000,
                 tolerance: float = 1e-6, association_threshold: float = 5.0):
        self.fc = fc
        self.c = 3e8
        self.lambda_c = lambda_c or (self.c / self.fc)
        self.num_antennas = num_antennas
        self.antenna_spacing = antenna_spacing or (self.lambda_c / 2)
        self.optimization_method = optimization_method
        self.max_iterations = max_iterations
        self.tolerance = tolerance
        self.association_threshold = association_threshold
        self.antenna_positions = np.arange(self.num_antennas) * self.antenna_spacing
        # not in the reference: the lattice search's knobs
        self.velocity_bounds = ((-50.0, 50.0), (-50.0, 50.0))          # :383, v_x and v_y
        self.lattice_points_per_period = 6.0
        self.polish_candidates = 48

    # ---- helpers
    @staticmethod
    def _device():
        return _device.pipeline(fc=77e9).device          # any pipeline: only the library handle and the device matter

    @staticmethod
    def _xy(targets) -> np.ndarray:
        rng = column_of(targets, 'range_m', float).reshape(-1)
        az = column_of(targets, 'azimuth_rad', float).reshape(-1)
        return np.stack([rng * np.cos(az), rng * np.sin(az)], axis=1).reshape(-1, 2)

    # ---- reference API
    def associate_targets_across_frames(self, current_targets: List[Dict], previous_targets: List[Dict]) -> List[Dict]:
        current, previous = records_of(current_targets), records_of(previous_targets)
        if not len(previous):
            logger.warning("No previous targets for association")
            return []
        if not len(current):
            return []
        dev = self._device()
        lib = _lib.load()
        cur = torch.from_numpy(np.ascontiguousarray(self._xy(current))).to(dev)
        prev = torch.from_numpy(np.ascontiguousarray(self._xy(previous))).to(dev)
        nc, npv = len(current), len(previous)
        n_cur = torch.tensor([nc], dtype=torch.int32, device=dev)
        n_prev = torch.tensor([npv], dtype=torch.int32, device=dev)
        match = torch.empty(nc, dtype=torch.int32, device=dev)
        dist = torch.empty(nc, dtype=torch.float64, device=dev)
        _lib.check(lib.rs_associate_targets(cur.data_ptr(), n_cur.data_ptr(), prev.data_ptr(), n_prev.data_ptr(),
                                            float(self.association_threshold), match.data_ptr(), dist.data_ptr(),
                                            1, nc, npv, torch.cuda.current_stream(dev).cuda_stream), "rs_associate_targets")
        match, dist = match.cpu().numpy(), dist.cpu().numpy()
        associations = []
        for i in np.nonzero(match >= 0)[0]:
            cur_t, prev_t = current[int(i)], previous[int(match[i])]
            associations.append({'current': cur_t, 'previous': prev_t, 'distance': float(dist[i]),
                                 'temporal_phase_diff': self._compute_temporal_phase_difference(cur_t, prev_t)})
        logger.info(f"Associated {len(associations)} targets across frames")
        return associations

    def _compute_temporal_phase_difference(self, current_target: Dict, previous_target: Dict) -> float:
        return np.angle(current_target['spatial_signature'][0] * np.conj(previous_target['spatial_signature'][0]))

    def compute_observed_phase_differences(self, target_associations: List[Dict]) -> np.ndarray:
        return np.array([a['temporal_phase_diff'] for a in target_associations])

    def compute_phase_difference_model(self, target_positions: np.ndarray, target_angles: np.ndarray,
                                       velocity: np.ndarray, angular_velocity: np.ndarray, dt: float) -> np.ndarray:
        pos = np.asarray(target_positions, dtype=float).reshape(-1, 3)
        ang = np.asarray(target_angles, dtype=float).reshape(-1, 2)
        d = np.stack([np.cos(ang[:, 1]) * np.cos(ang[:, 0]), np.cos(ang[:, 1]) * np.sin(ang[:, 0]), np.sin(ang[:, 1])], axis=1)
        rel = np.asarray(velocity, dtype=float)[None, :] + np.cross(np.asarray(angular_velocity, dtype=float)[None, :], pos)
        return (4 * np.pi * np.sum(rel * d, axis=1) * dt) / self.lambda_c

    def _costs(self, motions: np.ndarray, pos: np.ndarray, ang: np.ndarray, y: np.ndarray, dt: float) -> np.ndarray:
        """cost_function for a batch of motions [Q, 6] on the device."""
        dev = self._device()
        lib = _lib.load()
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)      # noqa: E731
        m, p, a, yy = t(motions.reshape(-1, 6)), t(pos.reshape(-1, 3)), t(ang.reshape(-1, 2)), t(y.reshape(-1))
        out = torch.empty(m.shape[0], dtype=torch.float64, device=dev)
        _lib.check(lib.rs_wrapped_cost(m.data_ptr(), p.data_ptr(), a.data_ptr(), yy.data_ptr(), int(yy.numel()),
                                       int(m.shape[0]), 4 * np.pi * dt / self.lambda_c, 0.01, 0.01, out.data_ptr(),
                                       torch.cuda.current_stream(dev).cuda_stream), "rs_wrapped_cost")
        return out.cpu().numpy()

    def cost_function(self, motion_params: np.ndarray, target_positions: np.ndarray, target_angles: np.ndarray,
                      observed_phases: np.ndarray, dt: float) -> float:
        return float(self._costs(np.asarray(motion_params, dtype=float), np.asarray(target_positions, dtype=float),
                                 np.asarray(target_angles, dtype=float), np.asarray(observed_phases, dtype=float), dt)[0])

    def get_smart_initial_guess(self, target_associations: List[Dict], dt: float) -> np.ndarray:
        if not target_associations:
            return np.array([0, 0, 0, 0, 0, 0])
        vel = []
        for a in target_associations:
            c, p = a['current'], a['previous']
            cp = np.array([c['range_m'] * np.cos(c['azimuth_rad']), c['range_m'] * np.sin(c['azimuth_rad']), 0])
            pp = np.array([p['range_m'] * np.cos(p['azimuth_rad']), p['range_m'] * np.sin(p['azimuth_rad']), 0])
            vel.append((cp - pp) / dt)
        med = np.median(np.array(vel), axis=0)
        return np.concatenate([np.append(-med[:2], 0), np.array([0, 0, 0])])

    # ---- the optimiser
    def _global_search(self, c: np.ndarray, s: np.ndarray, y: np.ndarray, k: float):
        """Global minimiser of sum wrap(y - k (vx c + vy s))^2 + 0.01 (vx^2 + vy^2) over the velocity box."""
        cand, cost, npts = wrapped_global_search(self._device(), c, s, y, k, self.velocity_bounds, reg_lattice=0.01,
                                                 reg_polish=0.01, centre=(0.0, 0.0),
                                                 points_per_period=self.lattice_points_per_period,
                                                 polish_candidates=self.polish_candidates)
        b = int(np.argmin(cost))
        return cand[b], float(cost[b]), npts

    def two_step_optimization(self, target_associations: List[Dict], dt: float,
                              initial_guess: Optional[np.ndarray] = None) -> Dict:
        if len(target_associations) < 3:
            logger.warning("Insufficient target associations for optimization")
            return {'success': False, 'message': 'Insufficient target associations'}
        rng = np.array([a['current']['range_m'] for a in target_associations], dtype=float)
        az = np.array([a['current']['azimuth_rad'] for a in target_associations], dtype=float)
        el = np.zeros_like(az)                                                   # "Assume ground level" (:342)
        target_positions = np.stack([rng * np.cos(el) * np.cos(az), rng * np.cos(el) * np.sin(az), rng * np.sin(el)], axis=1)
        target_angles = np.stack([az, el], axis=1)
        observed_phases = self.compute_observed_phase_differences(target_associations)
        k = 4 * np.pi * dt / self.lambda_c
        v2, f2, npts = self._global_search(np.cos(az), np.sin(az), observed_phases, k)
        velocity_est = np.array([v2[0], v2[1], 0.0])
        angular_velocity_est = np.zeros(3)
        full = np.concatenate([velocity_est, angular_velocity_est])
        cost_value = self.cost_function(full, target_positions, target_angles, observed_phases, dt)
        predicted_phases = self.compute_phase_difference_model(target_positions, target_angles, velocity_est,
                                                               angular_velocity_est, dt)
        residuals = observed_phases - predicted_phases
        residuals = np.arctan2(np.sin(residuals), np.cos(residuals))
        step1 = _Result(x=velocity_est.copy(), fun=cost_value, success=True, nfev=npts,
                        message='lattice search over the velocity box + Gauss-Newton polish')
        step2 = _Result(x=full.copy(), fun=cost_value, success=True, nfev=npts,
                        message='v_z and angular velocity do not enter the planar model: regulariser optimum 0')
        return {
            'success': True, 'velocity': velocity_est, 'angular_velocity': angular_velocity_est, 'cost': cost_value,
            'rmse': np.sqrt(np.mean(residuals ** 2)), 'max_residual': np.max(np.abs(residuals)), 'residuals': residuals,
            'predicted_phases': predicted_phases, 'observed_phases': observed_phases,
            'num_associations': len(target_associations), 'step1_result': step1, 'step2_result': step2,
        }

    def solve_velocity_with_association(self, current_targets: List[Dict], previous_targets: List[Dict],
                                        dt: float = 0.1) -> Dict:
        target_associations = self.associate_targets_across_frames(current_targets, previous_targets)
        if not target_associations:
            logger.warning("No target associations found")
            return {'success': False, 'message': 'No target associations'}
        return self.two_step_optimization(target_associations, dt)


def estimate_velocity_improved(current_angles_path: str, previous_angles_path: str, output_path: str,
                               radar_params: Dict = None, dt: float = 0.1) -> Dict:
    """velocity_solver_improved.py:503-558."""
    current_targets = records_of(np.load(current_angles_path, allow_pickle=True)['targets'])
    previous_targets = records_of(np.load(previous_angles_path, allow_pickle=True)['targets'])
    if radar_params is None:
        radar_params = {'fc': 77e9, 'lambda_c': 3e8 / 77e9, 'num_antennas': 8}
    solver = ImprovedVelocitySolver(**radar_params)
    results = solver.solve_velocity_with_association(current_targets, previous_targets, dt)
    np.savez(output_path, **results)
    return results
