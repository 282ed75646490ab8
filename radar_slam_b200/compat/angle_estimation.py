"""Drop-in for the reference's src/angle_estimation/angle_estimation.py (AngleEstimator,
extract_angles_from_rds, CLI) -- angle_estimation.py:23-417.

process_targets gathers every peak's unit-energy snapshot and evaluates the MUSIC / beamforming
pseudo-spectrum (fp64, returned per target like the reference) or the ESPRIT closed form on the GPU.
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .. import tables
from . import _device
from .lazy import LazyRecords, column_of, records_of

logger = logging.getLogger(__name__)


def _peaks_list(peak_info):
    """peak_info['peaks'] as a record sequence: LazyRecords (also unwrapped from the 0-d object array np.load
    returns), the reference's 1-D object array of dicts, or a list."""
    return records_of(peak_info['peaks'])


class AngleEstimator:
    """Angle of Arrival estimation for multi-channel FMCW radar (angle_estimation.py:23)."""

    def __init__(self, fc: float = 77e9, antenna_spacing: float = None, num_antennas: int = 8,
                 search_range: Tuple[float, float] = (-90, 90), search_resolution: float = 0.5):
        self.fc = fc
        self.c = 3e8
        self.lambda_c = self.c / self.fc
        self.antenna_spacing = antenna_spacing or (self.lambda_c / 2)
        self.num_antennas = num_antennas
        self.search_range = search_range
        self.search_resolution = search_resolution
        self.antenna_positions = np.arange(self.num_antennas) * self.antenna_spacing
        self.azimuth_grid = np.arange(search_range[0], search_range[1] + search_resolution, search_resolution)
        logger.info("Initialized angle estimator:")
        logger.info(f"  Antennas: {self.num_antennas}, spacing: {self.antenna_spacing*1000:.1f} mm")
        logger.info(f"  Search range: {search_range[0]}° to {search_range[1]}°")
        logger.info(f"  Search resolution: {search_resolution}°")

    # ---- device state
    def _pipe(self):
        return _device.pipeline(fc=self.fc, antenna_spacing=self.antenna_spacing, num_antennas=self.num_antennas,
                                search_range=tuple(self.search_range), search_resolution=self.search_resolution)

    def _steer128(self, pipe) -> torch.Tensor:
        # built from the object's own grid / positions so user edits of those attributes are honoured
        key = ("steer128", self.azimuth_grid.tobytes(), self.antenna_positions.tobytes(), self.lambda_c)
        if key not in pipe._tab:
            pipe._tab[key] = pipe._dev(tables.steering(self.azimuth_grid, self.antenna_positions, self.lambda_c))
        return pipe._tab[key]

    # ---- small helpers with the reference's semantics
    def extract_spatial_signature(self, rds: np.ndarray, range_bin: int, doppler_bin: int) -> np.ndarray:
        spatial_signature = rds[:, range_bin, doppler_bin]
        power = np.sum(np.abs(spatial_signature) ** 2)
        if power > 0:
            spatial_signature = spatial_signature / np.sqrt(power)
        return spatial_signature

    def generate_steering_vector(self, azimuth_deg: float) -> np.ndarray:
        phases = 2 * np.pi * self.antenna_positions * np.sin(np.radians(azimuth_deg)) / self.lambda_c
        return np.exp(1j * phases)

    def _spectrum(self, spatial_signature: np.ndarray, method: str) -> Tuple[np.ndarray, int]:
        pipe = self._pipe()
        sig = torch.from_numpy(np.ascontiguousarray(spatial_signature, dtype=np.complex128).reshape(1, -1)).to(pipe.device)
        spec, aidx = _device.spectra(pipe, sig, self._steer128(pipe), method)
        return spec[0].cpu().numpy(), int(aidx[0].item())

    def music_spectrum(self, spatial_signature: np.ndarray, num_sources: int = 1) -> np.ndarray:
        """angle_estimation.py:109-154.  The covariance is the rank-1 outer product of ONE snapshot, so for
        num_sources == 1 the noise-subspace projection is M - |a^H s|^2/|s|^2.  For num_sources > 1 the
        reference discards additional eigenvectors of the degenerate null space that LAPACK happens to return;
        that choice is not defined by the algorithm and is not reproduced."""
        if num_sources != 1:
            raise NotImplementedError("num_sources > 1 on a single-snapshot covariance depends on LAPACK's arbitrary "
                                      "null-space basis in the reference and is not supported")
        return self._spectrum(spatial_signature, "music")[0]

    def estimate_angle_music(self, spatial_signature: np.ndarray, num_sources: int = 1) -> Tuple[float, np.ndarray]:
        if num_sources != 1:
            self.music_spectrum(spatial_signature, num_sources)
        spec, idx = self._spectrum(spatial_signature, "music")
        return self.azimuth_grid[idx], spec

    def estimate_angle_esprit(self, spatial_signature: np.ndarray, num_sources: int = 1) -> float:
        try:
            pipe = self._pipe()
            sig = torch.from_numpy(np.ascontiguousarray(spatial_signature, dtype=np.complex128).reshape(1, -1)).to(pipe.device)
            if sig.shape[1] < 2:
                raise ValueError("need at least two antennas")
            scale = self.lambda_c / (2 * np.pi * self.antenna_spacing)
            return float(_device.esprit(pipe, sig, scale)[0].item())
        except Exception as e:          # angle_estimation.py:223-225
            logger.warning(f"ESPRIT failed: {e}")
            return 0.0

    def estimate_angle_beamforming(self, spatial_signature: np.ndarray) -> Tuple[float, np.ndarray]:
        spec, idx = self._spectrum(spatial_signature, "beamforming")
        return self.azimuth_grid[idx], spec

    def process_targets(self, rds: np.ndarray, peak_info: Dict, method: str = 'music') -> List[Dict]:
        """angle_estimation.py:253-309: one dict per peak (10 keys).  An unknown method is raised inside the
        reference's per-target try block, so it logs and yields an empty list; peaks whose bins fall outside
        the RDS are skipped the same way."""
        peaks = _peaks_list(peak_info)
        if method not in ('music', 'esprit', 'beamforming'):
            if len(peaks):
                logger.warning(f"Error processing target: Unknown method: {method}")
            logger.info(f"Processed 0 targets using {method}")
            return []
        rds = np.asarray(rds) if not isinstance(rds, np.ndarray) else rds
        A, R, D = rds.shape
        rb = column_of(peaks, 'range_bin', np.int64).reshape(-1)
        db = column_of(peaks, 'doppler_bin', np.int64).reshape(-1)
        ok = (rb >= -R) & (rb < R) & (db >= -D) & (db < D)
        for _ in range(int((~ok).sum())):
            logger.warning("Error processing target: index out of bounds")
        keep = np.nonzero(ok)[0]
        rbk, dbk = rb[keep] % R, db[keep] % D          # numpy negative indexing
        targets = []
        if len(keep):
            pipe = self._pipe()
            rds_dev = _device.rds_to_device(rds, pipe)
            sig_dev = _device.signatures(pipe, rds_dev, rbk, dbk)
            if method == 'esprit':
                scale = self.lambda_c / (2 * np.pi * self.antenna_spacing)
                angles = _device.esprit(pipe, sig_dev, scale).cpu().numpy()
                spec = None
            else:
                spec_dev, aidx = _device.spectra(pipe, sig_dev, self._steer128(pipe), method)
                angles = self.azimuth_grid[aidx.cpu().numpy()]
                spec = spec_dev.cpu().numpy()
            sigs = sig_dev.cpu().numpy()
            # angle_estimation.py:289-300: ten keys per target; columns now, dicts when somebody asks (compat/lazy.py)
            def peak_col(name):
                col = column_of(peaks, name)
                return col[keep] if isinstance(col, np.ndarray) and col.ndim >= 1 else col
            targets = LazyRecords({
                'range_m': peak_col('range_m'), 'doppler_hz': peak_col('doppler_hz'), 'power_db': peak_col('power_db'),
                'azimuth_deg': angles, 'azimuth_rad': np.radians(angles), 'antenna': peak_col('antenna'),
                'range_bin': peak_col('range_bin'), 'doppler_bin': peak_col('doppler_bin'),
                'spatial_signature': sigs, 'spectrum': spec,
            })
        logger.info(f"Processed {len(targets)} targets using {method}")
        return targets

    def visualize_angle_spectrum(self, targets: List[Dict], save_path: Optional[str] = None) -> None:
        if not targets:
            logger.warning("No targets to visualize")
            return
        import matplotlib.pyplot as plt
        fig, axes = plt.subplots(2, 2, figsize=(12, 10))
        angles = [t['azimuth_deg'] for t in targets]
        axes[0, 0].hist(angles, bins=20, alpha=0.7)
        axes[0, 0].set_xlabel('Azimuth Angle (degrees)')
        axes[0, 0].set_ylabel('Count')
        axes[0, 0].set_title('Angle Distribution')
        axes[0, 0].grid(True)
        ranges = [t['range_m'] for t in targets]
        powers = [t['power_db'] for t in targets]
        axes[0, 1].scatter(angles, ranges, c=powers, cmap='viridis', alpha=0.7)
        axes[0, 1].set_xlabel('Azimuth Angle (degrees)')
        axes[0, 1].set_ylabel('Range (m)')
        axes[0, 1].set_title('Range vs Angle')
        axes[0, 1].grid(True)
        axes[1, 0].scatter(angles, powers, alpha=0.7)
        axes[1, 0].set_xlabel('Azimuth Angle (degrees)')
        axes[1, 0].set_ylabel('Power (dB)')
        axes[1, 0].set_title('Power vs Angle')
        axes[1, 0].grid(True)
        if targets[0]['spectrum'] is not None:
            axes[1, 1].plot(self.azimuth_grid, 10 * np.log10(targets[0]['spectrum'] + 1e-12))
            axes[1, 1].set_xlabel('Azimuth Angle (degrees)')
            axes[1, 1].set_ylabel('MUSIC Spectrum (dB)')
            axes[1, 1].set_title('MUSIC Spectrum')
            axes[1, 1].grid(True)
        plt.tight_layout()
        if save_path:
            plt.savefig(save_path, dpi=150, bbox_inches='tight')
        plt.show()


def extract_angles_from_rds(rds_path: str, peak_info_path: str, output_path: str, method: str = 'music',
                            radar_params: Dict = None) -> Dict:
    """angle_estimation.py:368-417."""
    rds = np.load(rds_path)
    peak_info = dict(np.load(peak_info_path, allow_pickle=True))
    # 'peaks' comes back as the reference's 1-D object array of dicts or as the 0-d object array np.savez makes of a
    # LazyRecords (this package's own dechirp stage): unwrap before anything asks for its length
    peak_info['peaks'] = records_of(peak_info['peaks'])
    logger.info(f"Loaded RDS: {rds.shape}")
    logger.info(f"Found {len(peak_info['peaks'])} peaks")
    if radar_params is None:
        radar_params = {'fc': 77e9, 'antenna_spacing': 3e8 / (2 * 77e9), 'num_antennas': 8}
    estimator = AngleEstimator(**radar_params)
    targets = estimator.process_targets(rds, peak_info, method)
    np.savez(output_path, targets=targets, radar_params=radar_params)
    logger.info(f"Saved angle estimates for {len(targets)} targets")
    return {'num_targets': len(targets), 'method': method, 'targets': targets}


def main(argv=None):
    import argparse
    parser = argparse.ArgumentParser(description='Extract angles from RDS data')
    parser.add_argument('--rds', required=True, help='Path to RDS file')
    parser.add_argument('--peaks', required=True, help='Path to peak info file')
    parser.add_argument('--out', required=True, help='Output path for angles')
    parser.add_argument('--method', choices=['music', 'esprit', 'beamforming'], default='music',
                        help='Angle estimation method')
    args = parser.parse_args(argv)
    results = extract_angles_from_rds(args.rds, args.peaks, args.out, args.method)
    print(f"Angle extraction complete: {results}")
