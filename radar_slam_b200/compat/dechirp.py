"""Drop-in for the reference's src/radar_signal/dechirp.py (SignalPreprocessor, process_frame, CLI).

Same constructor, attributes, method names, argument meaning, return types and error behaviour
(dechirp.py:21-355); the frame-level work -- dechirp*window, DC removal, 2-D FFT + fftshift, dB power,
3x3 local maxima, threshold and range gate -- runs in libradarslam_b200.so on the GPU, in fp64 like the
reference (csrc/rs_legacy_f64.cu); the batched complex64 path with its fp64 recheck is radar_slam_b200.FramePipeline.
matplotlib is imported lazily so the module loads where it is not installed.
"""
from __future__ import annotations

import logging
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .. import tables
from . import _device
from .lazy import LazyRecords

logger = logging.getLogger(__name__)


class SignalPreprocessor:
    """Signal preprocessing for FMCW radar data (dechirp.py:21)."""

    def __init__(self, fc: float = 77e9, bandwidth: float = 1e9, chirp_duration: float = 40e-6, pri: float = 100e-6,
                 num_chirps: int = 64, sampling_rate: float = 10e6, window_type: str = 'hann', dc_removal: bool = True):
        self.fc = fc
        self.bandwidth = bandwidth
        self.chirp_duration = chirp_duration
        self.pri = pri
        self.num_chirps = num_chirps
        self.sampling_rate = sampling_rate
        self.window_type = window_type
        self.dc_removal = dc_removal
        self.c = 3e8
        self.lambda_c = self.c / self.fc
        self.samples_per_chirp = int(self.chirp_duration * self.sampling_rate)
        self.chirp_rate = self.bandwidth / self.chirp_duration
        self.range_resolution = self.c / (2 * self.bandwidth)
        self.velocity_resolution = self.lambda_c / (2 * self.num_chirps * self.pri)
        logger.info("Initialized signal preprocessor:")
        logger.info(f"  Range resolution: {self.range_resolution:.2f} m")
        logger.info(f"  Velocity resolution: {self.velocity_resolution:.2f} m/s")

    # ---- device pipeline for the current attribute values
    def _pipe(self):
        return _device.pipeline(fc=self.fc, bandwidth=self.bandwidth, chirp_duration=self.chirp_duration, pri=self.pri,
                                num_chirps=self.num_chirps, sampling_rate=self.sampling_rate,
                                window_type=self.window_type, dc_removal=bool(self.dc_removal))

    # ---- per-chirp helpers (API parity; the frame path below does not call them)
    def generate_reference_chirp(self) -> np.ndarray:
        return tables.reference_chirp(self.fc, self.chirp_rate, self.chirp_duration, self.samples_per_chirp)

    def apply_window(self, signal: np.ndarray, window_type: str = None) -> np.ndarray:
        if window_type is None:
            window_type = self.window_type
        return signal * tables.window(window_type, len(signal))

    def remove_dc(self, signal: np.ndarray) -> np.ndarray:
        return signal - np.mean(signal)

    def dechirp_signal(self, received_signal: np.ndarray, reference_chirp: Optional[np.ndarray] = None) -> np.ndarray:
        if reference_chirp is None:
            reference_chirp = self.generate_reference_chirp()
        return received_signal * np.conj(reference_chirp)

    def process_chirp(self, chirp_signal: np.ndarray, reference_chirp: Optional[np.ndarray] = None) -> np.ndarray:
        """Dechirp, window, remove DC for one chirp (dechirp.py:143-166), evaluated on the GPU in fp64."""
        if reference_chirp is None:
            reference_chirp = self.generate_reference_chirp()
        x = np.ascontiguousarray(chirp_signal, dtype=np.complex128)
        ref = np.ascontiguousarray(reference_chirp, dtype=np.complex128)
        if x.ndim != 1 or ref.shape != x.shape:
            raise ValueError(f"operands could not be broadcast together with shapes {x.shape} {ref.shape}")
        win = tables.window(self.window_type, len(x))
        pipe = self._pipe()
        dx, dr, dw = (torch.from_numpy(a).to(pipe.device) for a in (x, ref, win))
        out = torch.empty_like(dx)
        pipe._call("rs_process_chirps_f64", dx.data_ptr(), dr.data_ptr(), dw.data_ptr(), 1, len(x),
                   int(bool(self.dc_removal)), out.data_ptr(), pipe.stream)
        return out.cpu().numpy()

    # ---- frame path
    def generate_range_doppler_spectrum(self, frame_signals: np.ndarray,
                                        chirp_subset: Optional[Tuple[int, int]] = None) -> np.ndarray:
        """[num_antennas, num_chirps, samples] complex -> RDS [num_antennas, range_bins, doppler_bins]
        complex128, computed in fp64 (dechirp.py:168-213)."""
        frame_signals = np.asarray(frame_signals)
        num_antennas, num_chirps, samples_per_chirp = frame_signals.shape
        if samples_per_chirp != self.samples_per_chirp:
            raise ValueError(f"operands could not be broadcast together with shapes ({samples_per_chirp},) "
                             f"({self.samples_per_chirp},)")
        tables.window(self.window_type, 1)        # ValueError for an unknown window, like apply_window
        if chirp_subset is not None:
            start, end = chirp_subset
            start, end, _ = slice(start, end).indices(num_chirps)
            if end - start != chirp_subset[1] - chirp_subset[0] or end <= start:
                raise ValueError("could not broadcast input array: chirp_subset outside the frame")
            chirp_subset = (start, end)
        pipe = self._pipe()
        c0, c1 = chirp_subset if chirp_subset is not None else (0, num_chirps)
        Cu, S = c1 - c0, samples_per_chirp
        # the reference works in complex128 (dechirp.py:193-211): so does this method -- fp64 kernels on the very
        # array the caller passed, one frame per call (the batched complex64 path is FramePipeline)
        cube = torch.from_numpy(np.ascontiguousarray(frame_signals, dtype=np.complex128)).to(pipe.device)
        ref = torch.from_numpy(self.generate_reference_chirp().astype(np.complex128)).to(pipe.device)
        win = torch.from_numpy(np.ascontiguousarray(tables.window(self.window_type, S), dtype=np.float64)).to(pipe.device)
        tw_s = torch.from_numpy(tables.twiddles128(S)).to(pipe.device)
        tw_c = torch.from_numpy(tables.twiddles128(Cu)).to(pipe.device)
        rds = torch.empty((num_antennas, S, Cu), dtype=torch.complex128, device=pipe.device)
        pipe._call("rs_range_doppler_f64", cube.data_ptr(), ref.data_ptr(), win.data_ptr(), tw_s.data_ptr(),
                   tw_c.data_ptr(), rds.data_ptr(), num_antennas, num_chirps, c0, Cu, S, int(bool(self.dc_removal)),
                   pipe.stream)
        return rds.cpu().numpy()

    def extract_range_doppler_peaks(self, rds: np.ndarray, threshold_db: float = -20.0, min_range: float = 1.0,
                                    max_range: float = 200.0) -> Dict:
        """dechirp.py:215-278: {'peaks': [dict...], 'range_bins_m', 'doppler_bins_hz', 'power_spectrum_db'},
        decided in fp64 on the array passed in (wherever it came from)."""
        pipe = self._pipe()
        rds_dev = _device.rds_to_device(rds, pipe)
        A, R, D = rds_dev.shape
        gate = tables.range_gate(self.range_resolution, R, min_range, max_range)
        ant, rb, db, power_db = _device.detect_f64(pipe, rds_dev, gate, threshold_db)
        range_bins_m = tables.range_axis(self.range_resolution, R)
        doppler_bins_hz = tables.doppler_axis(self.sampling_rate, D)
        rm, dh, pw = range_bins_m[rb], doppler_bins_hz[db], power_db[ant, rb, db]
        # dechirp.py:265-272: one dict per peak; kept as columns and materialised on access (compat/lazy.py)
        peaks = LazyRecords({'antenna': ant, 'range_bin': rb, 'doppler_bin': db, 'range_m': rm, 'doppler_hz': dh,
                             'power_db': pw})
        return {'peaks': peaks, 'range_bins_m': range_bins_m, 'doppler_bins_hz': doppler_bins_hz,
                'power_spectrum_db': power_db}

    def visualize_rds(self, rds: np.ndarray, antenna_idx: int = 0, save_path: Optional[str] = None) -> None:
        import matplotlib.pyplot as plt
        power_db = 10 * np.log10(np.abs(rds[antenna_idx, :, :]) ** 2 + 1e-12)
        range_bins = np.linspace(0, self.range_resolution * rds.shape[1], rds.shape[1])
        doppler_bins = np.linspace(-self.sampling_rate / 2, self.sampling_rate / 2, rds.shape[2])
        plt.figure(figsize=(10, 6))
        plt.imshow(power_db, aspect='auto', origin='lower',
                   extent=[doppler_bins[0], doppler_bins[-1], range_bins[0], range_bins[-1]], cmap='jet')
        plt.colorbar(label='Power (dB)')
        plt.xlabel('Doppler Frequency (Hz)')
        plt.ylabel('Range (m)')
        plt.title(f'Range-Doppler Spectrum (Antenna {antenna_idx})')
        if save_path:
            plt.savefig(save_path, dpi=150, bbox_inches='tight')
        plt.show()


def process_frame(raw_signals_path: str, output_path: str, radar_params: Dict,
                  chirp_subset: Optional[Tuple[int, int]] = None) -> Dict:
    """dechirp.py:313-355: load .npy, RDS, peaks, save <out>.npy and <out>_peaks.npz."""
    raw_signals = np.load(raw_signals_path)
    logger.info(f"Loaded raw signals: {raw_signals.shape}")
    preprocessor = SignalPreprocessor(**radar_params)
    rds = preprocessor.generate_range_doppler_spectrum(raw_signals, chirp_subset)
    logger.info(f"Generated RDS: {rds.shape}")
    peak_info = preprocessor.extract_range_doppler_peaks(rds)
    logger.info(f"Found {len(peak_info['peaks'])} peaks")
    np.save(output_path, rds)
    np.savez(output_path.replace('.npy', '_peaks.npz'), **peak_info)
    return {'rds_shape': rds.shape, 'num_peaks': len(peak_info['peaks']), 'peak_info': peak_info}


def main(argv=None):
    import argparse
    parser = argparse.ArgumentParser(description='Process raw FMCW signals')
    parser.add_argument('--raw', required=True, help='Path to raw signals file')
    parser.add_argument('--out', required=True, help='Output path for RDS')
    parser.add_argument('--chirp-start', type=int, help='Start chirp index')
    parser.add_argument('--chirp-end', type=int, help='End chirp index')
    args = parser.parse_args(argv)
    radar_params = {'fc': 77e9, 'bandwidth': 1e9, 'chirp_duration': 40e-6, 'pri': 100e-6, 'num_chirps': 64,
                    'sampling_rate': 10e6}
    chirp_subset = None
    if args.chirp_start is not None and args.chirp_end is not None:
        chirp_subset = (args.chirp_start, args.chirp_end)
    results = process_frame(args.raw, args.out, radar_params, chirp_subset)
    print(f"Processing complete: {results}")
