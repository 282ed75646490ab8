"""The C-ABI library builds for sm_100a, loads without a GPU and exports exactly what
include/radar_slam_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "radar_slam_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rs_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    from radar_slam_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _declared()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.exported_symbols()) == declared
    _lib.load()
    assert _lib.load().rs_version() >= 100


def test_sm100a_cubin_present():
    import subprocess
    from radar_slam_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_argument_validation_without_gpu():
    """Entry points validate before touching CUDA: bad sizes return RS_EINVAL with a message."""
    from radar_slam_b200 import _lib
    lib = _lib.load()
    rc = lib.rs_range_fft(None, None, None, None, 1, 1, 1, 0, 1, 8, 1, None)
    assert rc == -1 and b"null" in lib.rs_last_error()
    tr = ctypes.c_int()
    td = ctypes.c_int()
    nt = ctypes.c_int()
    assert lib.rs_detect_tiling(256, 128, 8, ctypes.byref(tr), ctypes.byref(td), ctypes.byref(nt)) == 0
    assert tr.value * td.value > 0 and nt.value == ((256 + tr.value - 1) // tr.value) * ((128 + td.value - 1) // td.value)
    assert lib.rs_detect_tiling(0, 128, 8, None, None, None) == -1


def test_no_product_import_of_oracle():
    """The product must never reach into oracle/ (parity would be void)."""
    bad = []
    for base in ("radar_slam_b200", "src", "scripts"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith(".py"):
                    txt = open(os.path.join(dp, fn)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad
