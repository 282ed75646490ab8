"""The reference's UNMODIFIED scripts/run_ego_motion_pipeline.py resolves the hot-path modules from this repo
when the repo precedes the reference on PYTHONPATH (namespace-package shadowing, SURVEY.md 8b / probe p11), while
everything outside the hot path (pose integration, evaluation, the simulator) still comes from the reference.
Import only -- no GPU work.  Skipped where /root/reference does not exist (the GPU box)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


def _run(code, tmp_path):
    env = dict(os.environ)
    shims = os.path.join(ROOT, "tests", "shims")
    try:
        import matplotlib  # noqa: F401
        extra = []
    except ImportError:
        extra = [shims]
    env["PYTHONPATH"] = os.pathsep.join([ROOT] + extra)
    return subprocess.run([sys.executable, "-c", code], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=300)


def test_unmodified_pipeline_script_imports_this_repo(tmp_path):
    code = (
        "import sys; sys.path.insert(0, '/root/reference/scripts'); sys.argv=['run_ego_motion_pipeline.py']\n"
        "import run_ego_motion_pipeline as m\n"
        "print(m.SignalPreprocessor.__module__, m.AngleEstimator.__module__, m.VelocitySolver.__module__,\n"
        "      m.PoseIntegrator.__module__, m.FMCWRadarSimulator.__module__)\n"
        "import inspect; print(inspect.getsourcefile(m))\n"
        "p = m.EgoMotionPipeline('seq', '/nonexistent', 'out', max_frames=1)\n"
        "print(sorted(p.radar_params))\n"
    )
    r = _run(code, tmp_path)
    assert r.returncode == 0, r.stderr[-2000:]
    mods = r.stdout.splitlines()[0].split()
    assert mods[0] == "radar_slam_b200.compat.dechirp"
    assert mods[1] == "radar_slam_b200.compat.angle_estimation"
    assert mods[2] == "radar_slam_b200.compat.velocity_solver"
    assert mods[3] == "src.pose_integration.pose_integration"          # still the reference's
    assert mods[4] == "simulate_raw"
    assert r.stdout.splitlines()[1].startswith("/root/reference/scripts/")


def test_reference_tests_import_this_repo(tmp_path):
    """tests/test_phase2_simplified.py:22 imports src.algorithms.robust_angle_estimation; tests/test_synth_raw.py
    imports src.radar_signal.dechirp."""
    code = (
        "import sys; sys.path.append('/root/reference')\n"
        "from src.algorithms.robust_angle_estimation import RobustAngleEstimator\n"
        "from src.robust_angle_estimation import RobustAngleEstimator as R2\n"
        "from src.radar_signal.dechirp import SignalPreprocessor\n"
        "from src.algorithms.velocity_solver_improved import ImprovedVelocitySolver\n"
        "from src.algorithms.advanced_velocity_optimization import AdvancedVelocityOptimizer\n"
        "print(RobustAngleEstimator.__module__, R2 is RobustAngleEstimator, SignalPreprocessor.__module__,\n"
        "      ImprovedVelocitySolver.__module__, AdvancedVelocityOptimizer.__module__)\n"
    )
    r = _run(code, tmp_path)
    assert r.returncode == 0, r.stderr[-2000:]
    out = r.stdout.split()
    assert out[0] == "radar_slam_b200.compat.robust_angle_estimation" and out[1] == "True"
    assert out[2] == "radar_slam_b200.compat.dechirp"
    assert out[3] == "radar_slam_b200.compat.velocity_solver_improved"
    assert out[4] == "radar_slam_b200.compat.advanced_velocity_optimization"


def test_constructor_signatures_match_reference():
    import inspect
    sys.path.insert(0, ROOT)
    from oracle import ref_import
    ref = ref_import.load()
    from radar_slam_b200.compat import dechirp, angle_estimation, robust_angle_estimation, velocity_solver
    pairs = [(ref.SignalPreprocessor, dechirp.SignalPreprocessor), (ref.AngleEstimator, angle_estimation.AngleEstimator),
             (ref.RobustAngleEstimator, robust_angle_estimation.RobustAngleEstimator),
             (ref.VelocitySolver, velocity_solver.VelocitySolver)]
    for rc, mc in pairs:
        for name, fn in inspect.getmembers(rc, predicate=inspect.isfunction):
            if name.startswith("_") and name != "__init__":
                continue
            assert hasattr(mc, name), f"{mc.__name__} lacks {name}"
            rs, ms = inspect.signature(fn), inspect.signature(getattr(mc, name))
            assert list(rs.parameters) == list(ms.parameters), (mc.__name__, name)
            for k in rs.parameters:
                assert rs.parameters[k].default == ms.parameters[k].default or \
                    rs.parameters[k].default is inspect._empty, (mc.__name__, name, k)
    for rm, mm, fns in [(ref.dechirp, dechirp, ["process_frame"]), (ref.angle, angle_estimation, ["extract_angles_from_rds"]),
                        (ref.robust, robust_angle_estimation, ["extract_angles_robust"]),
                        (ref.velocity, velocity_solver, ["estimate_velocity_from_angles"])]:
        for fn in fns:
            assert list(inspect.signature(getattr(rm, fn)).parameters) == list(inspect.signature(getattr(mm, fn)).parameters)


def test_advanced_optimizer_signatures_match_reference():
    """src/algorithms/advanced_velocity_optimization.py keeps the reference's class, methods and defaults."""
    import inspect
    sys.path.insert(0, ROOT)
    from oracle import ref_import
    ref_import.load()
    ref = ref_import._load(os.path.join(REF, "src/algorithms/advanced_velocity_optimization.py"), "_rsref_advanced_sig")
    from radar_slam_b200.compat import advanced_velocity_optimization as mine
    for name, fn in inspect.getmembers(ref.AdvancedVelocityOptimizer, predicate=inspect.isfunction):
        if name.startswith("__") and name != "__init__":
            continue
        assert hasattr(mine.AdvancedVelocityOptimizer, name), name
        rs, ms = inspect.signature(fn), inspect.signature(getattr(mine.AdvancedVelocityOptimizer, name))
        assert list(rs.parameters) == list(ms.parameters), name
        for k in rs.parameters:
            assert rs.parameters[k].default == ms.parameters[k].default, (name, k)
    assert list(inspect.signature(ref.optimize_velocity_advanced).parameters) == \
        list(inspect.signature(mine.optimize_velocity_advanced).parameters)


def test_real_time_shell_signatures_match_reference():
    """src/core/real_time_processor.py keeps the reference's classes, methods and defaults."""
    import importlib.util
    import inspect
    sys.path.insert(0, ROOT)
    spec = importlib.util.spec_from_file_location("ref_real_time_processor", os.path.join(REF, "src/core/real_time_processor.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from radar_slam_b200.compat import real_time_processor as mine
    for cls in ("FrameBuffer", "ParallelTargetProcessor", "RealTimeProcessor", "RealTimeVelocityEstimator"):
        rc, mc = getattr(ref, cls), getattr(mine, cls)
        for name, fn in inspect.getmembers(rc, predicate=inspect.isfunction):
            if name.startswith("__") and name != "__init__":
                continue
            assert hasattr(mc, name), f"{cls} lacks {name}"
            rs, ms = inspect.signature(fn), inspect.signature(getattr(mc, name))
            assert list(rs.parameters) == list(ms.parameters), (cls, name)
            for k in rs.parameters:
                assert rs.parameters[k].default == ms.parameters[k].default, (cls, name, k)
    assert [f.name for f in ref.ProcessingFrame.__dataclass_fields__.values()] == \
        [f.name for f in mine.ProcessingFrame.__dataclass_fields__.values()]
    assert list(inspect.signature(ref.create_real_time_estimator).parameters) == \
        list(inspect.signature(mine.create_real_time_estimator).parameters)
