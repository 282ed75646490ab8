"""ctypes binding of libradarslam_b200.so (include/radar_slam_b200.h).

The library is the product: if it cannot be loaded this module raises -- there is no CPU or
PyTorch fallback.  torch supplies device memory and the stream only; every signature is plain
pointers and sizes.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

RS_METHOD_MUSIC, RS_METHOD_BEAMFORMING, RS_METHOD_ESPRIT = 0, 1, 2
RS_RECHECK_FRAME_CAP = 256
RS_TIE_LIST_CAP = 32
RS_ANGLES_WS_BYTES = 32 << 20
RS_FLAG_TIE, RS_FLAG_NEARMAX, RS_FLAG_GUARD, RS_FLAG_FIXED, RS_FLAG_DROPPED, RS_FLAG_DETFIXED = 1, 2, 4, 8, 16, 32
METHODS = {"music": RS_METHOD_MUSIC, "beamforming": RS_METHOD_BEAMFORMING, "esprit": RS_METHOD_ESPRIT}

_vp, _i, _f, _d = C.c_void_p, C.c_int, C.c_float, C.c_double

_SIGNATURES = {
    "rs_version": (_i, []),
    "rs_last_error": (C.c_char_p, []),
    "rs_detect_tiling": (_i, [_i, _i, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "rs_range_fft": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "rs_doppler_fft": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "rs_range_doppler_fft": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "rs_fft2d_ws_max_clusters": (_i, []),
    "rs_detect": (_i, [_vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "rs_fused_detect_ws_bytes": (C.c_longlong, [_i, _i]),
    "rs_range_doppler_detect": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _f, _f,
                                     _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "rs_detection_power": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "rs_angles": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _f, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                       _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "rs_velocity_partials": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "rs_velocity_from_partials": (_i, [_vp, _i, _i, _d, _d, _vp, _vp, _vp]),
    "rs_music_covariance": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "rs_recheck_detections_f64": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _d, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i,
                                       _vp, _vp, _vp, _vp, _vp, _vp]),
    "rs_recheck_angles_f64": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _d, _vp, _vp, _vp, _vp, _vp,
                                   _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp,
                                   _vp, _vp]),
    "rs_velocity_ls": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _d, _d, _i, _d, _vp, _i, _i, _i, _vp]),
    "rs_rds_to_reference_layout": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "rs_rds_from_reference_layout": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "rs_signatures_f64": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _vp]),
    "rs_spectra_f64": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "rs_power_db_f64": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "rs_process_chirps_f64": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "rs_range_doppler_f64": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "rs_detect_f64": (_i, [_vp, _vp, _d, _vp, _vp, _vp, _vp, C.c_longlong, _vp, _i, _i, _i, _vp]),
    "rs_signatures_c128": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp]),
    "rs_esprit_f64": (_i, [_vp, _i, _i, _d, _vp, _vp]),
    "rs_velocity_ls6": (_i, [_vp, _vp, _vp, _i, _d, _vp, _vp, _i, _vp, _vp, _vp]),
    "rs_associate_targets": (_i, [_vp, _vp, _vp, _vp, _d, _vp, _vp, _i, _i, _i, _vp]),
    "rs_wrapped_cost": (_i, [_vp, _vp, _vp, _vp, _i, _i, _d, _d, _d, _vp, _vp]),
    "rs_regularized_cost": (_i, [_vp, _vp, _vp, _vp, _i, _i, _d, _d, _d, _d, _vp, _vp, _vp]),
    "rs_wrapped_gn_polish": (_i, [_vp, _vp, _vp, _i, _d, _d, _d, _d, _d, _d, _d, _d, _vp, _vp, _i, _i, _vp]),
    "rs_wrapped_lattice_tiles": (_i, [C.c_longlong, C.c_longlong, C.POINTER(_i), C.POINTER(_i)]),
    "rs_wrapped_lattice_search": (_i, [_vp, _vp, _vp, _i, _d, _d, _d, C.c_longlong, C.c_longlong, _d, _vp, _vp, _vp, _vp]),
    "rs_synthesize_frames": (_i, [_vp, _vp, _i, _d, _d, _d, _d, _vp, _d, C.c_ulonglong, C.c_longlong, _vp, _vp,
                                  _i, _i, _i, _i, _vp]),
    "rs_robust_confidence_f64": (_i, [_vp, _vp, _vp, _d, _i, _i, _vp, _vp]),
}

_lib = None


class RadarSlamError(RuntimeError):
    pass


def exported_symbols():
    return sorted(_SIGNATURES)


def load():
    """Load (building first if the .so is absent and nvcc is available).  Raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("RADAR_SLAM_B200_LIB", _build.LIB_PATH)
    if not os.path.exists(path):
        path = _build.build()
    lib = C.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().rs_last_error().decode(errors="replace")
        raise RadarSlamError(f"{what or 'libradarslam_b200'} failed ({rc}): {msg}")


def ptr(t) -> int:
    """Device (or host) address of a torch tensor / None."""
    return 0 if t is None else t.data_ptr()


def detect_tiling(R: int, D: int, A: int):
    tr, td, nt = _i(), _i(), _i()
    check(load().rs_detect_tiling(R, D, A, C.byref(tr), C.byref(td), C.byref(nt)), "rs_detect_tiling")
    return tr.value, td.value, nt.value
