"""Drop-in for the reference's src/robust_angle_estimation.py (and its byte-identical copy
src/algorithms/robust_angle_estimation.py): RobustAngleEstimator, extract_angles_robust, CLI
(robust_angle_estimation.py:23-570).

Per frame the GPU does the snapshot gather, the beamforming scan with first-index argmax and the
confidence metric for the selected top-K peaks; the power filter / stable sort / top-K selection and
the temporal smoothing are sequential host state exactly like the reference.
"""
from __future__ import annotations

import logging
import time
from collections import deque
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .. import tables
from . import _device
from .angle_estimation import _peaks_list
from .lazy import column_of, records_of

logger = logging.getLogger(__name__)


class RobustAngleEstimator:
    def __init__(self, fc: float = 77e9, antenna_spacing: float = None, num_antennas: int = 8,
                 search_range: Tuple[float, float] = (-90, 90), search_resolution: float = 1.0,
                 temporal_window: int = 5, confidence_threshold: float = 0.7, smoothing_factor: float = 0.3,
                 max_targets: int = 100):
        self.fc = fc
        self.c = 3e8
        self.lambda_c = self.c / self.fc
        self.antenna_spacing = antenna_spacing or (self.lambda_c / 2)
        self.num_antennas = num_antennas
        self.search_range = search_range
        self.search_resolution = search_resolution
        self.temporal_window = temporal_window
        self.confidence_threshold = confidence_threshold
        self.smoothing_factor = smoothing_factor
        self.max_targets = max_targets
        self.antenna_positions = np.arange(self.num_antennas) * self.antenna_spacing
        self.azimuth_grid = np.arange(search_range[0], search_range[1] + search_resolution, search_resolution)
        self.angle_history = {}
        self.confidence_history = {}
        self.target_counter = 0
        logger.info("Initialized robust angle estimator:")
        logger.info(f"  Temporal window: {temporal_window}")
        logger.info(f"  Confidence threshold: {confidence_threshold}")
        logger.info(f"  Smoothing factor: {smoothing_factor}")
        logger.info(f"  Max targets: {max_targets}")

    # ---- device helpers
    def _pipe(self):
        return _device.pipeline(fc=self.fc, antenna_spacing=self.antenna_spacing, num_antennas=self.num_antennas,
                                search_range=tuple(self.search_range), search_resolution=self.search_resolution)

    def _steer128(self, pipe) -> torch.Tensor:
        key = ("steer128", self.azimuth_grid.tobytes(), self.antenna_positions.tobytes(), self.lambda_c)
        if key not in pipe._tab:
            pipe._tab[key] = pipe._dev(tables.steering(self.azimuth_grid, self.antenna_positions, self.lambda_c))
        return pipe._tab[key]

    def _confidence_batch(self, pipe, sig_dev: torch.Tensor, angles_deg: np.ndarray) -> np.ndarray:
        n, A = sig_dev.shape
        ang = torch.from_numpy(np.ascontiguousarray(angles_deg, dtype=np.float64)).to(pipe.device)
        pos = pipe._dev(np.ascontiguousarray(self.antenna_positions, dtype=np.float64))
        out = torch.empty((n,), dtype=torch.float64, device=pipe.device)
        pipe._call("rs_robust_confidence_f64", sig_dev.data_ptr(), ang.data_ptr(), pos.data_ptr(), float(self.lambda_c),
                   n, A, out.data_ptr(), pipe.stream)
        return out.cpu().numpy()

    def _estimate_batch(self, sig: np.ndarray):
        """beamforming spectra, initial angles and confidences for snapshots [n, A] (host complex128)."""
        pipe = self._pipe()
        sig_dev = torch.from_numpy(np.ascontiguousarray(sig, dtype=np.complex128)).to(pipe.device)
        spec, aidx = _device.spectra(pipe, sig_dev, self._steer128(pipe), "beamforming")
        init = self.azimuth_grid[aidx.cpu().numpy()]
        conf = self._confidence_batch(pipe, sig_dev, init)
        return spec.cpu().numpy(), init, conf

    # ---- reference API
    def generate_steering_vector(self, azimuth_deg: float) -> np.ndarray:
        phases = 2 * np.pi * self.antenna_positions * np.sin(np.radians(azimuth_deg)) / self.lambda_c
        return np.exp(1j * phases)

    def compute_angle_confidence(self, spatial_signature: np.ndarray, estimated_angle: float) -> float:
        pipe = self._pipe()
        sig_dev = torch.from_numpy(np.ascontiguousarray(spatial_signature, dtype=np.complex128).reshape(1, -1)).to(pipe.device)
        return float(self._confidence_batch(pipe, sig_dev, np.array([float(estimated_angle)]))[0])

    def detect_multipath_interference(self, spatial_signature: np.ndarray) -> Dict:
        """robust_angle_estimation.py:140-218.  The covariance is the rank-1 outer product of one snapshot: its
        spectrum is {|s|^2, 0, ..., 0}.  In the reference the 'geometric' and arithmetic noise means are the same
        expression (:177-179), so the MDL data term vanishes and the penalty selects one source; the remaining
        figures (snr_ratio, condition_number) are ratios against LAPACK's ~1e-17 rounding residue of the zero
        eigenvalues and are reported here at their exact values (infinite)."""
        s = np.asarray(spatial_signature)
        N = len(s)
        eigenvals = np.zeros(N)
        eigenvals[0] = float(np.sum(np.abs(s) ** 2))
        return {'num_sources': 1, 'snr_ratio': float('inf'), 'condition_number': float('inf'),
                'eigenvalues': eigenvals, 'is_multipath': False, 'interference_level': 0.0}

    def apply_temporal_smoothing(self, target_id: str, new_angle: float, new_confidence: float) -> Tuple[float, float]:
        """robust_angle_estimation.py:274-330 -- sequential host state."""
        if target_id not in self.angle_history:
            self.angle_history[target_id] = deque(maxlen=self.temporal_window)
            self.confidence_history[target_id] = deque(maxlen=self.temporal_window)
        self.angle_history[target_id].append(new_angle)
        self.confidence_history[target_id].append(new_confidence)
        if len(self.angle_history[target_id]) >= 2:
            angles = np.array(self.angle_history[target_id])
            confidences = np.array(self.confidence_history[target_id])
            total = np.sum(confidences)
            weights = confidences / total if total > 0 else np.ones_like(confidences) / len(confidences)
            angles_rad = np.radians(angles)
            mean_cos = np.sum(weights * np.cos(angles_rad))
            mean_sin = np.sum(weights * np.sin(angles_rad))
            smoothed_angle = np.degrees(np.arctan2(mean_sin, mean_cos))
            prev_angle = self.angle_history[target_id][-2]
            smoothed_angle = self.smoothing_factor * smoothed_angle + (1 - self.smoothing_factor) * prev_angle
            smoothed_confidence = np.mean(confidences)
        else:
            smoothed_angle = new_angle
            smoothed_confidence = new_confidence
        return smoothed_angle, smoothed_confidence

    def _finish(self, sig, spec, initial_angle, confidence, target_id) -> Dict:
        interference_analysis = self.detect_multipath_interference(sig)
        if target_id is not None:
            smoothed_angle, smoothed_confidence = self.apply_temporal_smoothing(target_id, initial_angle, confidence)
        else:
            smoothed_angle, smoothed_confidence = initial_angle, confidence
        is_reliable = (smoothed_confidence >= self.confidence_threshold and not interference_analysis['is_multipath'])
        return {'angle_deg': smoothed_angle, 'angle_rad': np.radians(smoothed_angle), 'confidence': smoothed_confidence,
                'is_reliable': is_reliable, 'interference_analysis': interference_analysis, 'spectrum': spec,
                'initial_angle': initial_angle, 'smoothing_applied': target_id is not None}

    def estimate_angle_robust(self, spatial_signature: np.ndarray, target_id: str = None) -> Dict:
        spec, init, conf = self._estimate_batch(np.asarray(spatial_signature).reshape(1, -1))
        return self._finish(np.asarray(spatial_signature), spec[0], init[0], float(conf[0]), target_id)

    def process_targets_robust(self, rds: np.ndarray, peak_info: Dict, frame_timestamp: float = None) -> List[Dict]:
        """robust_angle_estimation.py:346-411."""
        peaks = _peaks_list(peak_info)
        # robust_angle_estimation.py:362-370: power filter, stable descending sort, top max_targets -- on the power
        # column, so only the selected peaks are ever materialised as dicts
        power = column_of(peaks, 'power_db', float).reshape(-1) if len(peaks) else np.zeros(0)
        cand = np.nonzero(power > -25.0)[0]
        cand = cand[np.argsort(-power[cand], kind='stable')][:self.max_targets]
        filtered_peaks = [peaks[int(i)] for i in cand]
        targets: List[Dict] = []
        if filtered_peaks:
            A, R, D = rds.shape
            pipe = self._pipe()
            rds_dev = _device.rds_to_device(rds, pipe)
            rb = np.array([int(p['range_bin']) for p in filtered_peaks], dtype=np.int64)
            db = np.array([int(p['doppler_bin']) for p in filtered_peaks], dtype=np.int64)
            ok = (rb >= -R) & (rb < R) & (db >= -D) & (db < D)
            sig_dev = _device.signatures(pipe, rds_dev, np.where(ok, rb % R, 0), np.where(ok, db % D, 0))
            spec_dev, aidx = _device.spectra(pipe, sig_dev, self._steer128(pipe), "beamforming")
            init = self.azimuth_grid[aidx.cpu().numpy()]
            conf = self._confidence_batch(pipe, sig_dev, init)
            sigs = sig_dev.cpu().numpy()
            spec = spec_dev.cpu().numpy()
            for i, peak in enumerate(filtered_peaks):
                if not ok[i]:
                    logger.warning("Error processing target: index out of bounds")
                    continue
                target_id = f"target_{peak['range_bin']}_{peak['doppler_bin']}"
                res = self._finish(sigs[i], spec[i], init[i], float(conf[i]), target_id)
                if res['is_reliable']:
                    targets.append({
                        'range_m': peak['range_m'], 'doppler_hz': peak['doppler_hz'], 'power_db': peak['power_db'],
                        'azimuth_deg': res['angle_deg'], 'azimuth_rad': res['angle_rad'],
                        'confidence': res['confidence'], 'is_reliable': res['is_reliable'],
                        'interference_analysis': res['interference_analysis'], 'antenna': peak['antenna'],
                        'range_bin': peak['range_bin'], 'doppler_bin': peak['doppler_bin'],
                        'spatial_signature': sigs[i], 'target_id': target_id,
                        'timestamp': frame_timestamp or time.time(),
                    })
        logger.info(f"Processed {len(targets)} reliable targets (filtered from {len(filtered_peaks)})")
        return targets

    def get_target_statistics(self) -> Dict:
        total_targets = len(self.angle_history)
        active_targets = sum(1 for h in self.angle_history.values() if len(h) > 0)
        all_confidences = []
        for confidences in self.confidence_history.values():
            all_confidences.extend(confidences)
        avg_confidence = np.mean(all_confidences) if all_confidences else 0.0
        return {'total_targets_tracked': total_targets, 'active_targets': active_targets,
                'average_confidence': avg_confidence, 'temporal_window_size': self.temporal_window,
                'confidence_threshold': self.confidence_threshold}

    def visualize_angle_quality(self, targets: List[Dict], save_path: Optional[str] = None) -> None:
        if not targets:
            logger.warning("No targets to visualize")
            return
        import matplotlib.pyplot as plt
        fig, axes = plt.subplots(2, 2, figsize=(15, 10))
        angles = [t['azimuth_deg'] for t in targets]
        confidences = [t['confidence'] for t in targets]
        axes[0, 0].hist(angles, bins=20, alpha=0.7, edgecolor='black')
        axes[0, 0].set_title('Angle Distribution')
        axes[0, 1].hist(confidences, bins=20, alpha=0.7, edgecolor='black', color='green')
        axes[0, 1].axvline(self.confidence_threshold, color='red', linestyle='--')
        axes[0, 1].set_title('Confidence Distribution')
        axes[1, 0].scatter(angles, confidences, alpha=0.7, s=50)
        axes[1, 0].axhline(self.confidence_threshold, color='red', linestyle='--')
        axes[1, 0].set_title('Angle vs Confidence')
        levels = [t['interference_analysis']['interference_level'] for t in targets]
        axes[1, 1].hist(levels, bins=20, alpha=0.7, edgecolor='black', color='orange')
        axes[1, 1].set_title('Interference Analysis')
        plt.tight_layout()
        if save_path:
            plt.savefig(save_path, dpi=150, bbox_inches='tight')
        plt.show()


def extract_angles_robust(rds_path: str, peak_info_path: str, output_path: str, radar_params: Dict = None,
                          temporal_window: int = 5, confidence_threshold: float = 0.7) -> Dict:
    """robust_angle_estimation.py:508-570."""
    rds = np.load(rds_path)
    peak_info = dict(np.load(peak_info_path, allow_pickle=True))
    # 'peaks' comes back as the reference's 1-D object array of dicts or as the 0-d object array np.savez makes of a
    # LazyRecords (this package's own dechirp stage): unwrap before anything asks for its length
    peak_info['peaks'] = records_of(peak_info['peaks'])
    logger.info(f"Loaded RDS: {rds.shape}")
    logger.info(f"Found {len(peak_info['peaks'])} peaks")
    if radar_params is None:
        radar_params = {'fc': 77e9, 'antenna_spacing': 3e8 / (2 * 77e9), 'num_antennas': 8}
    estimator = RobustAngleEstimator(**radar_params, temporal_window=temporal_window,
                                     confidence_threshold=confidence_threshold)
    targets = estimator.process_targets_robust(rds, peak_info)
    stats = estimator.get_target_statistics()
    np.savez(output_path, targets=targets, radar_params=radar_params, statistics=stats)
    logger.info(f"Saved robust angle estimates for {len(targets)} targets")
    logger.info(f"Statistics: {stats}")
    return {'num_targets': len(targets), 'statistics': stats, 'targets': targets}


def main(argv=None):
    import argparse
    parser = argparse.ArgumentParser(description='Extract angles with robust estimation')
    parser.add_argument('--rds', required=True, help='Path to RDS file')
    parser.add_argument('--peaks', required=True, help='Path to peak info file')
    parser.add_argument('--out', required=True, help='Output path for angles')
    parser.add_argument('--temporal-window', type=int, default=5, help='Temporal window size')
    parser.add_argument('--confidence-threshold', type=float, default=0.7, help='Confidence threshold')
    args = parser.parse_args(argv)
    results = extract_angles_robust(args.rds, args.peaks, args.out, temporal_window=args.temporal_window,
                                    confidence_threshold=args.confidence_threshold)
    print(f"Robust angle extraction complete: {results}")
