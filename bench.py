#!/usr/bin/env python
"""Benchmark of the radar-slam per-frame hot path on B200 (contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload auto|configs1|configs4] [--impl reference]

A "step" is one pass of the hot path (dechirp -> range FFT -> Doppler FFT -> peak detection -> MUSIC -> least-squares
ego-velocity) over one batch of synthetic frames.

  N = 1 (headline)   BASELINE.json configs[1]: 1k frames of 256 samples x 128 chirps x 8 channels, MUSIC on a 1 degree
                     grid, reference-default noise and threshold, the 1.95 GiB cube resident in HBM (>> L2).
                     The same line carries `configs4_n1`: the N > 1 workload measured on this one GPU, and
                     `configs2_n1` / `configs3_n1`: BASELINE configs[2] (512 x 256 x 192, ESPRIT) and the configs[3] style
                     scene (threshold 31 dB, 3 Huber iterations) on this GPU.
  N > 1              BASELINE.json configs[4]: ONE 65 536-frame sequence of 256 x 128 x 16 cubes, STRONG-scaled: rank r
                     owns the contiguous frame block frame_block(r, N, 65536), walks it in chunks through a resident pool
                     of synthetic frames (inputs resident in HBM when the timed region starts; pool >> L2), the solve
                     writes into the rank's slot of the all-gather buffer and one NCCL all_gather_into_tensor of the
                     [65536, 8] velocity rows ends the step.  `configs1_weak` carries the round-1 weak-scaling figure.

The JSON line carries: value (frames/s, inputs resident in HBM), e2e (the same through FramePipeline.process_host with
pinned HOST buffers, H2D + D2H + the all-gather inside the timed region, next to a copy-only ceiling measured with the
same buffers), roofline of the dominant kernel and of every stage, cpu_baseline (the oracle port timed on this box's
host cores), a >= 1 s sustained run, clocks, gpu_launches.
`--impl reference` times the reference algorithm's CPU port (oracle/) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/s dechirp->range-Doppler->MUSIC->LS ego-velocity"
WORKLOAD = "configs[1]: 1k-frame synthetic batch per GPU, 256 samples x 128 chirps x 8 channels, MUSIC 1 deg grid, noise_power 0.01, threshold -20 dB"
WORKLOAD4 = ("configs[4]: one 65536-frame synthetic sequence, 256 samples x 128 chirps x 16 channels, frame-sharded by contiguous "
             "block, MUSIC 1 deg grid, noise_power 0.01, threshold -20 dB, NCCL all-gather of the velocity rows")
# kernels behind one C-ABI call (for gpu_launches): the 2-D transform alone is the persistent cluster kernel + the side
# kernel; with the detection fused in it is the persistent kernel + the compaction kernel (no side kernels by default)
KERNELS_PER_CALL = {"rs_range_doppler_fft": 2, "rs_range_doppler_detect": 2, "rs_detect": 1, "rs_angles": 1, "rs_recheck_detections_f64": 1,
                    "rs_recheck_angles_f64": 3, "rs_velocity_from_partials": 1, "rs_velocity_partials": 1, "rs_velocity_ls": 1,
                    "rs_range_fft": 1, "rs_doppler_fft": 1}


def workload_name(args):
    if (args.frames, args.samples, args.chirps, args.antennas, args.method, args.grid_res, args.threshold_db, args.irls) == \
            (1000, 256, 128, 8, "music", 1.0, -20.0, 0):
        return WORKLOAD
    return (f"non-default: {args.frames} frames per GPU, {args.samples} samples x {args.chirps} chirps x {args.antennas} channels, "
            f"{args.method} ({args.grid_res} deg grid), noise_power 0.01, threshold {args.threshold_db} dB, "
            f"{args.irls} Huber iterations")


def radar_config(args):
    from radar_slam_b200 import RadarConfig
    return RadarConfig(fc=77e9, bandwidth=1e9, chirp_duration=args.samples / 10e6, pri=100e-6, num_chirps=args.chirps,
                       sampling_rate=10e6, num_antennas=args.antennas, search_resolution=args.grid_res,
                       method=args.method, threshold_db=args.threshold_db, recheck=not args.no_recheck,
                       fft_eps=args.fft_eps, irls_iters=args.irls, huber_delta=args.huber)


def oracle_params(args):
    from oracle import radar_oracle as orc
    return orc.RadarParams(chirp_duration=args.samples / 10e6, num_chirps=args.chirps, num_antennas=args.antennas)


# ---------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------- CPU arms
def _oracle_worker(payload):
    frames, pdict, res, thr, blas_threads, method = payload
    from oracle import radar_oracle as orc
    p = orc.RadarParams(**pdict)
    tim = {}
    out = []

    def run():
        for fr in frames:
            r = orc.process_frame(fr.astype(np.complex128), p, method, res, thr, timings=tim)
            out.append((r["velocity"].get("velocity", np.zeros(3))[:2], len(r["peaks"]["antenna"])))

    if blas_threads:
        # one process per core: keep each worker's BLAS / OpenMP pools at one thread, or they oversubscribe the box
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=blas_threads):
            run()
    else:
        run()
    return out, tim


def cpu_oracle_rate(frames: np.ndarray, args, procs: int, pool=None):
    """Frames/s of the oracle port (vectorised numpy restatement of the reference) on `procs` processes.
    `pool`: a multiprocessing pool to reuse (the reference arm keeps one alive over all its steps)."""
    from dataclasses import asdict
    p = oracle_params(args)
    pdict = asdict(p)
    parts = [frames[i::procs] for i in range(procs)]
    parts = [x for x in parts if len(x)]
    t0 = time.perf_counter()
    if len(parts) == 1 and pool is None:
        res = [_oracle_worker((parts[0], pdict, args.grid_res, args.threshold_db, 1 if procs == 1 else 0, args.method))]
    else:
        import multiprocessing as mp
        payload = [(x, pdict, args.grid_res, args.threshold_db, 1, args.method) for x in parts]
        if pool is not None:
            res = pool.map(_oracle_worker, payload)
        else:
            with mp.get_context("fork").Pool(len(parts)) as own:
                res = own.map(_oracle_worker, payload)
    dt = time.perf_counter() - t0
    tim = {}
    ndet = []
    for out, t in res:
        for k, v in t.items():
            tim[k] = tim.get(k, 0.0) + v
        ndet += [n for _, n in out]
    return len(frames) / dt, dt, tim, float(np.mean(ndet)), res


def host_frames(args, n: int, seed: int) -> np.ndarray:
    """n frames of the benchmark workload generated on the HOST (numpy) for the CPU arms."""
    from radar_slam_b200 import synth
    cfg = radar_config(args)
    sig = synth.scatterer_term(cfg).astype(np.complex64)
    rs = np.random.RandomState(seed)
    A, C, S = args.antennas, args.chirps, args.samples
    out = np.empty((n, A, C, S), dtype=np.complex64)
    for i in range(n):
        noise = np.sqrt(0.01) * (rs.randn(A, C, S) + 1j * rs.randn(A, C, S))
        out[i] = (sig[:, None, :] + noise).astype(np.complex64)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    wl = resolve_workload(args, int(os.environ.get("WORLD_SIZE", "1")))
    if wl == "configs4":
        args.antennas = 16
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, args.ref_procs or cores))
    per_step = args.ref_frames_per_step or 2 * procs          # fixed sample per step: ~1-3 s of CPU work per step
    frames = host_frames(args, per_step, 4242)
    # one worker pool for the whole run: forking a pool inside the timed loop cost up to 60 % of a step in round 1
    with mp.get_context("fork").Pool(procs) as pool:
        for _ in range(args.warmup):
            cpu_oracle_rate(frames[:procs], args, procs, pool)
        t0 = time.perf_counter()
        ndet = 0.0
        for _ in range(args.steps):
            _, _, _, ndet, _ = cpu_oracle_rate(frames, args, procs, pool)
        dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "strong" if wl == "configs4" else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD4 if wl == "configs4" else workload_name(args), "samples": args.samples,
                   "chirps": args.chirps, "channels": args.antennas, "frames_per_step": per_step,
                   "detections_per_frame": ndet},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": procs, "kind": "port",
                         "sample": f"{per_step} frames/step of the same workload through oracle.radar_oracle.process_frame "
                                   f"(vectorised fp64 numpy port of the reference; the reference itself is Python and cannot "
                                   f"travel to this box), {procs} worker processes kept alive over all steps"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def resolve_workload(args, world: int) -> str:
    if args.workload != "auto":
        return args.workload
    return "configs1" if world == 1 else "configs4"


# ---------------------------------------------------------------------------------- GPU arm
def stage_bytes(name, F, A, C, S, n_det):
    cells = F * A * C * S
    if name in ("rs_range_fft", "rs_doppler_fft", "rs_range_doppler_fft"):
        return 16 * cells                                   # read c64 + write c64 per cell (the fused 2-D kernel: cube in, RDS out)
    if name == "rs_range_doppler_detect":
        # cube in, RDS out; the detection rides on the plane in shared memory: its traffic is the hit masks (written by the
        # FFT kernel, read by the compaction) and the lists (key + flag + leader per detection)
        return 16 * cells + 2 * (cells // 8 + cells // 1024) + 9 * n_det
    if name == "rs_detect":
        return 8 * cells + 9 * n_det                        # read RDS once; key + power + flag per detection
    if name == "rs_angles":
        return (8 * A + 4 + 2 + 12) * n_det                 # snapshot gather + key + flag r/w + aidx/adeg/phase
    if name == "rs_velocity_ls":
        return 9 * n_det + 64 * F                           # aidx + phase + flag per detection; one row per frame
    if name == "rs_recheck_angles_f64":
        return 8 * cells                                    # exact snapshots need every raw plane of the frame once
    if name == "rs_velocity_from_partials":
        return 64 * 16 * F + 64 * F                         # per-segment sums in, one row per frame out
    return 0


def _barrier(world, dev):
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def _max_over_ranks(x: float, world, dev) -> float:
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _time_steps(step, steps, warmup, world, dev, finish=None):
    """W untimed warm-up steps, then exactly `steps` steps between barrier + synchronize on both sides; CUDA events on the
    launching stream, max over ranks.  `finish` joins whatever the steps left on other streams (inside the timed region).
    Returns total milliseconds."""
    import torch
    for _ in range(warmup):
        step()
    if finish is not None:
        finish()
    _barrier(world, dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    if finish is not None:
        finish()
    e1.record()
    _barrier(world, dev)
    return _max_over_ranks(e0.elapsed_time(e1), world, dev)


def _e2e(pipe, host, vel_slot, gather, chunk, steps, world, dev, frames_this_rank):
    """End to end through FramePipeline.process_host with pinned host buffers: H2D of every chunk, kernels, the all-gather
    of the velocity rows and the D2H of the result inside the timed region; then the same buffers / chunking / streams
    with the kernels left out (copy-only ceiling of this platform), all ranks concurrently.  Wall clock, max over ranks."""
    import torch
    vel_host = torch.empty((frames_this_rank, 8), dtype=torch.float64, pin_memory=True)

    def run(copy_only):
        pipe.process_host(host, chunk_frames=chunk, vel_dev=vel_slot, vel_host=vel_host, copy_only=copy_only)
        if not copy_only:
            gather.gather()
            torch.cuda.synchronize(dev)

    out = {}
    for name, copy_only in (("e2e", False), ("ceiling", True)):
        for _ in range(2):
            run(copy_only)
        _barrier(world, dev)
        t0 = time.perf_counter()
        for _ in range(steps):
            run(copy_only)
        torch.cuda.synchronize(dev)
        out[name] = _max_over_ranks(time.perf_counter() - t0, world, dev) / steps
    return out, vel_host


def measure_configs4(args, world, rank, dev, steps, warmup):
    """BASELINE configs[4]: one 65 536-frame 256 x 128 x 16 sequence, strong-scaled by contiguous frame block."""
    import copy
    import torch
    from radar_slam_b200 import FramePipeline, synth
    from radar_slam_b200.sharding import VelocityGather, frame_block
    a4 = copy.copy(args)
    a4.antennas, a4.samples, a4.chirps = 16, 256, 128
    cfg = radar_config(a4)
    pipe = FramePipeline(cfg, device=str(dev))
    total = args.total_frames
    lo, hi = frame_block(rank, world, total)
    block = hi - lo
    P = min(block, args.pool_frames)
    A, C, S = 16, 128, 256
    # resident pool of this rank's first P frames of the sequence (frame k depends only on (seed, k)); the block is walked
    # through the pool chunk by chunk, so the inputs are in HBM when the timed region starts and far exceed the L2
    pool = synth.synth_cubes(cfg, P, seed=1234, first_frame=lo, device=dev)
    gather = VelocityGather(total, dev)
    vel = gather.slot()

    def step():
        # the steps are software pipelined: the fp64 recheck + solve of one launch set run on a side stream beside the FFT /
        # angle kernels of the next one, the all-gather follows the last solve of the step on that stream
        for c0 in range(0, block, P):
            n = min(P, block - c0)
            pipe.process(pool[:n], chunk_frames=args.chunk, vel_out=vel[c0:c0 + n], join=False,
                         after_solve=gather.gather if c0 + P >= block else None)

    pipe.launches = 0
    ms_total = _time_steps(step, steps, warmup, world, dev, finish=pipe.join)
    calls = pipe.launches // (steps + warmup)
    # ---- end to end on a bounded sample of the block (PCIe bound: a whole block would take seconds per step)
    n_e2e = min(block, args.e2e_frames4)
    host = torch.empty((n_e2e, A, C, S), dtype=torch.complex64, pin_memory=True)
    host.copy_(pool[:min(P, n_e2e)].repeat(((n_e2e + P - 1) // P), 1, 1, 1)[:n_e2e] if n_e2e > P else pool[:n_e2e])
    torch.cuda.synchronize(dev)
    t, _ = _e2e(pipe, host, vel, gather, args.host_chunk, max(1, min(steps, 3)), world, dev, n_e2e)
    bytes_step = n_e2e * A * C * S * 8
    res = {
        "workload": WORKLOAD4, "total_frames": total, "frames_this_rank": block, "resident_pool_frames": P,
        "value": total * steps / (ms_total / 1e3), "unit": "frames/s", "ms_per_step": ms_total / steps, "scaling": "strong",
        "stage_calls_per_step": calls,
        "e2e": {"value": world * n_e2e / t["e2e"], "unit": "frames/s", "frames_per_step_per_gpu": n_e2e,
                "h2d_bytes_per_step": bytes_step, "d2h_bytes_per_step": n_e2e * 64,
                "h2d_ceiling_gbs_per_gpu": bytes_step / t["ceiling"] / 1e9,
                "ceiling_frames_per_s": world * n_e2e / t["ceiling"],
                "frac_of_ceiling": t["ceiling"] / t["e2e"],
                "note": "ceiling = the same pinned buffers, chunking and streams with the kernels left out, all ranks "
                        "concurrently; the all-gather is inside the end-to-end region"},
    }
    del pool, host
    torch.cuda.empty_cache()
    return res


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from radar_slam_b200 import FramePipeline, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    from radar_slam_b200.sharding import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local) if not args.no_numa_bind else {"device": local, "numa_node": "unbound"}
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = resolve_workload(args, world)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    c4 = None
    if wl == "configs4" or (world == 1 and not args.no_configs4 and args.workload == "auto"):
        steps4 = args.steps if wl == "configs4" else max(2, min(args.steps, 3))
        c4 = measure_configs4(args, world, rank, dev, steps4, args.warmup if wl == "configs4" else 3)
    c1 = measure_configs1(args, world, rank, dev, numa, detailed=(wl == "configs1"))
    # N = 1, default arguments: BASELINE configs[2] and configs[3] ride along as side objects (same code path, their
    # own shapes; a failure there is reported in the object and never costs the headline line)
    side = {}
    if world == 1 and args.workload == "auto" and not args.no_side_configs and workload_name(args) == WORKLOAD:
        import argparse
        import gc
        for key, over in (("configs2_n1", dict(samples=512, chirps=256, antennas=192, method="esprit", frames=8, chunk=8,
                                                 e2e_frames=8, host_chunk=4, steps=max(3, min(args.steps, 5)))),
                          ("configs3_n1", dict(threshold_db=31.0, irls=3, steps=max(3, min(args.steps, 5))))):
            gc.collect()
            torch.cuda.empty_cache()
            try:
                a2 = argparse.Namespace(**{**vars(args), **over})
                r2 = measure_configs1(a2, world, rank, dev, numa, detailed=False)
                side[key] = {k: r2[k] for k in ("value", "unit", "ms_per_step", "steps", "scaling", "e2e", "config", "gpu_launches")}
            except Exception as exc:                              # noqa: BLE001 -- reported, not raised
                side[key] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    clocks = sampler.stop() if rank == 0 else None

    if rank == 0:
        if wl == "configs1":
            line = c1
            if c4 is not None:
                line["configs4_n1" if world == 1 else "configs4"] = c4
            line.update(side)
        else:
            line = {
                "metric": METRIC, "value": c4["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": c4["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD4, "total_frames": c4["total_frames"], "samples": 256, "chirps": 128,
                           "channels": 16, "frames_this_rank": c4["frames_this_rank"],
                           "resident_pool_frames": c4["resident_pool_frames"], "chunk_frames": args.chunk,
                           "cache": "inputs larger than L2 (resident pool of %.2f GiB per GPU, walked block by block)"
                                    % (c4["resident_pool_frames"] * 16 * 128 * 256 * 8 / 2 ** 30),
                           "parallelism": f"one {c4['total_frames']}-frame sequence, contiguous frame blocks over {world} GPU(s), "
                                          f"all-gather of [F,8] velocity rows",
                           "host_binding": numa},
                "roofline": c1["roofline"], "cpu_baseline": c1["cpu_baseline"], "e2e": c4["e2e"],
                "gpu_launches": int(c4["stage_calls_per_step"] * args.steps * 8 / 5),    # 8 kernels behind the 5 stage calls
                "configs1_weak": {k: c1[k] for k in ("value", "ms_per_step", "scaling", "e2e", "config")},
            }
        line["clocks"] = clocks
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_configs1(args, world, rank, dev, numa, detailed=True):
    """BASELINE configs[1] per GPU (weak scaling at N > 1); with `detailed` also the per-stage roofline, the CPU baseline
    and the sustained run.  Returns the JSON line (dict) on every rank."""
    import torch
    from radar_slam_b200 import FramePipeline, synth
    from radar_slam_b200.sharding import VelocityGather
    cfg = radar_config(args)
    pipe = FramePipeline(cfg, device=str(dev))
    F, A, C, S = args.frames, args.antennas, args.chirps, args.samples

    cube = synth.synth_cubes(cfg, F, seed=1234, first_frame=rank * F, device=dev)
    gather = VelocityGather(world * F, dev)       # all-gather target; this rank's slot is the solve's output
    vel = gather.slot()

    def step():
        # Steps are software pipelined (FramePipeline.process(join=False)): the fp64 recheck + solve of step k run on a side
        # stream beside the FFT / angle kernels of step k + 1 (two workspace sets); the NCCL all_gather_into_tensor of the
        # [F, 8] rows (no-op at N=1) follows the solve on that stream.  The last step is joined inside the timed region.
        pipe.process(cube, chunk_frames=args.chunk, vel_out=vel, join=False, after_solve=gather.gather)

    pipe.launches = 0
    pipe.call_counts = {}
    ms_total = _time_steps(step, args.steps, args.warmup, world, dev, finish=pipe.join)
    launches = sum(KERNELS_PER_CALL.get(k, 1) * v for k, v in pipe.call_counts.items()) * args.steps // (args.steps + args.warmup)

    # ---- sustained: the same step for >= 1 s (thermally / power settled), reported next to the K-step figure
    sustained = None
    if detailed and args.sustain_s > 0:
        n_sus = max(args.steps, int(args.sustain_s * 1e3 / (ms_total / args.steps)) + 1)
        ms_sus = _time_steps(step, n_sus, 0, world, dev, finish=pipe.join)
        sustained = {"steps": n_sus, "seconds": ms_sus / 1e3, "value": world * F * n_sus / (ms_sus / 1e3), "unit": "frames/s"}

    # ---- end to end through the public API with HOST buffers (pinned): H2D + kernels + all-gather + D2H in the timed region
    n_e2e = min(F, args.e2e_frames)
    host = torch.empty((n_e2e, A, C, S), dtype=torch.complex64, pin_memory=True)
    host.copy_(cube[:n_e2e])
    torch.cuda.synchronize(dev)
    e2e_steps = max(1, min(args.steps, 3))
    t_e2e, vel_host = _e2e(pipe, host, vel, gather, args.host_chunk, e2e_steps, world, dev, n_e2e)
    e2e_value = world * n_e2e / t_e2e["e2e"]
    pipe.process(cube, chunk_frames=args.chunk, vel_out=vel)
    pipe.process_host(host, chunk_frames=args.host_chunk, vel_dev=gather.buf.new_empty((n_e2e, 8)), vel_host=vel_host)
    torch.cuda.synchronize(dev)
    assert torch.equal(vel_host.to(dev), vel[:n_e2e]), "host path and device path disagree"
    bytes_step = n_e2e * A * C * S * 8
    e2e = {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": bytes_step, "d2h_bytes_per_step": n_e2e * 64,
           "frames_per_step": n_e2e, "h2d_ceiling_gbs_per_gpu": bytes_step / t_e2e["ceiling"] / 1e9,
           "ceiling_frames_per_s": world * n_e2e / t_e2e["ceiling"], "frac_of_ceiling": t_e2e["ceiling"] / t_e2e["e2e"],
           "note": "ceiling = the same pinned buffers, chunking and streams with the kernels left out, all ranks concurrently; "
                   "the all-gather is inside the end-to-end region"}

    value = world * F * args.steps / (ms_total / 1e3)
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "frames_per_gpu": F, "samples": S, "chirps": C, "channels": A,
                   "chunk_frames": args.chunk, "host_chunk_frames": args.host_chunk,
                   "cache": "inputs larger than L2 (%.2f GiB cube per GPU)" % (cube.numel() * 8 / 2 ** 30),
                   "parallelism": f"frames sharded over {world} GPU(s), all-gather of [F,8] velocity rows",
                   "host_binding": numa},
        "e2e": e2e, "gpu_launches": launches,
    }
    if sustained is not None:
        line["sustained"] = sustained
    if not detailed or rank != 0:
        line.setdefault("roofline", None)
        line.setdefault("cpu_baseline", None)
        return line

    # ---- per-stage device times (CUDA events on the launching stream) for the roofline: averaged over several passes
    os.environ["RS_NO_OVERLAP"] = "1"             # stage times one after the other (the timed run above overlaps the recheck)
    stage_ms, stage_n = {}, {}
    passes = max(1, args.profile_passes)
    for _ in range(passes):
        pipe.profile = []
        pipe.process(cube, chunk_frames=args.chunk, vel_out=vel)
        torch.cuda.synchronize(dev)
        for name, a, b in pipe.profile:
            stage_ms[name] = stage_ms.get(name, 0.0) + a.elapsed_time(b) / passes
            stage_n[name] = stage_n.get(name, 0) + 1
    stage_n = {k: v // passes for k, v in stage_n.items()}
    os.environ.pop("RS_NO_OVERLAP")
    pipe.profile = None
    # detections per frame (for the algorithmic bytes of the list-driven stages)
    rds = pipe.range_doppler(cube[: min(F, 64)])
    det = pipe.detect(rds)
    n_det_frame = float(det.per_frame_counts().double().mean().item())
    pipe.angles(rds, det)
    vm = det.valid_mask().reshape(-1)
    fl = det.flags[: vm.numel()][vm]
    flagged = {name: float(((fl & bit) != 0).sum().item()) / det.F for name, bit in (("tie", 1), ("nearmax", 2), ("guard", 4))}
    overflow = int(det.overflow.sum().item())
    recheck_stats = None
    if cfg.recheck:
        ds = pipe.recheck_detections(cube[: det.F], det).cpu().numpy().astype(float) / det.F
        pipe.angles(rds, det)
        as_ = pipe.recheck_angles(cube[: det.F], rds, det).cpu().numpy().astype(float) / det.F
        recheck_stats = {"detections_per_frame": dict(zip(("rechecked", "dropped", "promoted", "unresolved"), ds.tolist())),
                         "angles_per_frame": dict(zip(("rechecked", "index_changed", "fp64_snapshot", "unresolved"), as_.tolist()))}
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    stages = []
    tot = sum(stage_ms.values())
    for name, msv in stage_ms.items():
        by = stage_bytes(name, F, A, C, S, n_det_frame * F)
        stages.append({"kernel": name, "ms_per_step": msv, "share": msv / tot, "launches": stage_n[name],
                       "alg_bytes_per_step": by, "achieved_gbs": by / msv / 1e6, "frac_hbm": by / msv / 1e6 / peak})
    dom = max(stages, key=lambda s: s["ms_per_step"])
    per_launch_bytes = dom["alg_bytes_per_step"] / dom["launches"]
    # measured DRAM traffic of the same kernels from the committed ncu capture (per frame of this shape), if present
    traffic = None
    import glob
    import re
    tfiles = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_v*_traffic.json")),
                    key=lambda q: [int(x) for x in re.findall(r"\d+", os.path.basename(q))])
    tpath = tfiles[-1] if tfiles else ""                 # the newest committed capture
    if tpath and (A, C, S) == (8, 128, 256):
        per_frame = json.load(open(tpath))["dram_bytes_per_frame"]
        for st in stages:
            if st["kernel"] in per_frame:
                st["ncu_dram_bytes_per_step"] = per_frame[st["kernel"]] * F
        if dom["kernel"] in per_frame:
            traffic = per_frame[dom["kernel"]] * F / dom["launches"]
    roofline = {"kernel": dom["kernel"], "bound": "hbm", "achieved": dom["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": dom["frac_hbm"], "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_launch": per_launch_bytes, "avg_launch_ms": dom["ms_per_step"] / dom["launches"],
                "stages": stages}
    fft = [st for st in stages if st["kernel"] in ("rs_range_fft", "rs_doppler_fft", "rs_range_doppler_fft")]
    ms_fft = sum(st["ms_per_step"] for st in fft)
    if not fft and (C, S) == (128, 256):
        # the timed path runs the 2-D FFT with the detection fused into it (rs_range_doppler_detect); the transform alone
        # (rs_range_doppler_fft: the same kernel without the detection) is timed here for the FFT-stage roofline
        out = pipe._buf("rds0", (F, S, A, C), torch.complex64)
        for _ in range(2):
            pipe.range_doppler(cube, out=out)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(passes):
            pipe.range_doppler(cube, out=out)
        ev[1].record()
        torch.cuda.synchronize(dev)
        ms_fft = ev[0].elapsed_time(ev[1]) / passes
    if ms_fft > 0:
        roofline["fft_stages"] = {
            "ms_per_step": ms_fft, "alg_bytes_per_step": 16 * F * A * C * S,
            "frac_hbm_2d": 16 * F * A * C * S / ms_fft / 1e6 / peak,
            "note": "the whole 2-D transform (dechirp, window, range FFT, Doppler FFT, both shifts) against its 16 B/cell "
                    "(cube in, RDS out); one cluster kernel when the plane fits distributed shared memory"}
    ang = next((st for st in stages if st["kernel"] == "rs_angles"), None)
    if ang is not None:
        # the angle stage is not HBM bound: report its arithmetic rate next to its (contractual) HBM fraction
        G = len(pipe._angle_tables(A)["grid"])
        ap = 2 if A <= 2 else 4 if A <= 4 else 8 if A <= 8 else 16
        cells = float(det.nlead[: det.F * det.ntiles].sum().item()) / det.F * F
        flops = cells * (G * 2 * (ap - 1) + 8 * ap * ap)
        roofline["angle_scan"] = {
            "distinct_cells_per_step": cells, "scan_tflops_fp32_equivalent": flops / ang["ms_per_step"] / 1e9,
            "note": "per-cell MUSIC grid scan = [cells x lags] . [lags x grid pairs] on the 5th-generation tensor cores "
                    "(tcgen05.mma 128x32x16 into TMEM, fp16 operands split hi + lo, fp32 accumulation, persistent CTAs); its "
                    "two-level argmax / runner-up tracking is instruction-issue bound, not HBM bound"}
    fd = next((st for st in stages if st["kernel"] == "rs_range_doppler_detect"), None)
    if fd is not None:
        # what the two-stage path (rs_range_doppler_fft + rs_detect) needs for the same result: 16 + 8 B/cell + the lists
        two_stage = 24 * F * A * C * S + 9 * n_det_frame * F
        fd["two_stage_alg_bytes_per_step"] = two_stage
        fd["frac_hbm_of_two_stage_bytes"] = two_stage / fd["ms_per_step"] / 1e6 / peak
        if (C, S) == (128, 256) and A % 8 == 0:
            # the stage is two kernels (persistent FFT + detection kernel, then the mask compaction): the first one alone,
            # the compaction left out (RS_FD_NO_COMPACT), against cube in + RDS out + the masks it writes
            out = pipe._buf("rds0", (F, S, A, C), torch.complex64)
            os.environ["RS_FD_NO_COMPACT"] = "1"
            try:
                for _ in range(2):
                    pipe.range_doppler_detect(cube, out=out, workspace="0", defer_power=True)
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                ev[0].record()
                for _ in range(passes):
                    pipe.range_doppler_detect(cube, out=out, workspace="0", defer_power=True)
                ev[1].record()
                torch.cuda.synchronize(dev)
                ms_k = ev[0].elapsed_time(ev[1]) / passes
                kb = (16 + 0.27) * F * A * C * S
                fd["fft_detect_kernel_alone"] = {"ms_per_step": ms_k, "alg_bytes_per_step": kb, "frac_hbm": kb / ms_k / 1e6 / peak}
            finally:
                os.environ.pop("RS_FD_NO_COMPACT", None)
    if dom["kernel"] == "rs_range_doppler_detect":
        roofline["note"] = ("dominant stage is the fused 2-D FFT + detection (persistent warp-specialised 4-CTA clusters fed by TMA; "
                            "the |X|^2 local-maximum test runs on the plane in distributed shared memory and leaves hit masks, a "
                            "small kernel compacts them into the lists): the RDS is written once and not read by a detection "
                            "pass; `achieved` counts cube in + RDS out + masks + lists")
    elif dom["kernel"] == "rs_range_doppler_fft":
        roofline["note"] = ("dominant kernel is the fused 2-D FFT (persistent warp-specialised 4-CTA clusters fed by TMA + the side "
                            "kernel on the stranded SMs): HBM traffic at the algorithmic minimum (16 B/cell, plane held in "
                            "distributed shared memory between the passes)")
    elif dom["kernel"] == "rs_angles":
        roofline["note"] = ("dominant kernel by time is the per-cell MUSIC grid scan (see angle_scan): ALU-issue bound, not HBM "
                            "bound -- its HBM fraction is reported because the contract asks for one; the HBM-bound stages are "
                            "in fft_stages / stages")

    # ---- CPU baseline: the oracle port on a bounded sample of the SAME frames, one core
    n_cpu = args.cpu_frames
    if n_cpu > 0:
        sample = cube[:n_cpu].cpu().numpy()
        rate, dt, tim, ndet_cpu, res = cpu_oracle_rate(sample, args, 1)
        v_cpu = np.stack([v for v, _ in res[0][0]])
        v_gpu = vel[:n_cpu, :2].cpu().numpy()
        parity = float(np.abs(v_cpu - v_gpu).max())
    else:                                         # exploratory runs of the big configs: no CPU arm
        rate, dt, tim, ndet_cpu, parity = None, 0.0, {}, 0.0, None

    line["config"].update({"grid_points": len(pipe._angle_tables(A)["grid"]), "detections_per_frame": n_det_frame,
                           "detection_overflow": overflow, "undecided_in_fp32_per_frame": flagged,
                           "fp64_recheck": recheck_stats, "profile_passes": passes})
    line["roofline"] = roofline
    line["cpu_baseline"] = {"value": rate, "unit": "frames/s", "cores": 1, "kind": "port",
                            "sample": f"{n_cpu} frames of the same batch through the oracle port in {dt:.1f}s "
                                      f"(stage seconds {json.dumps({k: round(v, 2) for k, v in tim.items()})}); "
                                      f"{ndet_cpu:.0f} detections/frame",
                            "max_abs_velocity_diff_vs_gpu": parity}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--samples", type=int, default=256)
    ap.add_argument("--chirps", type=int, default=128)
    ap.add_argument("--antennas", type=int, default=8)
    ap.add_argument("--grid-res", type=float, default=1.0)
    ap.add_argument("--method", default="music", choices=["music", "beamforming", "esprit"])
    ap.add_argument("--irls", type=int, default=0, help="Huber reweighting iterations of the velocity solve (configs[3])")
    ap.add_argument("--huber", type=float, default=1.0)
    ap.add_argument("--threshold-db", type=float, default=-20.0)
    ap.add_argument("--chunk", type=int, default=1000, help="frames per launch set, device-resident path")
    ap.add_argument("--host-chunk", type=int, default=32, help="frames per H2D chunk, host-buffer path")
    ap.add_argument("--no-recheck", action="store_true", help="skip the fp64 recheck of flagged decisions (fp32 path only)")
    ap.add_argument("--fft-eps", type=float, default=4e-7, help="error bound of the fp32 FFT used by the recheck, in rms units")
    ap.add_argument("--e2e-frames", type=int, default=1000)
    ap.add_argument("--cpu-frames", type=int, default=24)
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin each rank to its GPU's NUMA node")
    ap.add_argument("--workload", default="auto", choices=["auto", "configs1", "configs4"],
                    help="auto: configs[1] at N = 1 (plus configs[4] on one GPU as `configs4_n1`), configs[4] strong-scaled at N > 1")
    ap.add_argument("--total-frames", type=int, default=65536, help="length of the configs[4] sequence")
    ap.add_argument("--pool-frames", type=int, default=512, help="configs[4]: resident synthetic frames per GPU (2 GiB at 512)")
    ap.add_argument("--e2e-frames4", type=int, default=1024, help="configs[4]: frames per GPU of the end-to-end sample")
    ap.add_argument("--no-configs4", action="store_true", help="N = 1: skip the configs[4] side measurement")
    ap.add_argument("--no-side-configs", action="store_true", help="N = 1: skip the configs[2] / configs[3] side measurements")
    ap.add_argument("--sustain-s", type=float, default=1.0, help="length of the sustained run (0: off)")
    ap.add_argument("--profile-passes", type=int, default=5, help="passes averaged for the per-stage times")
    ap.add_argument("--ref-procs", type=int, default=0)
    ap.add_argument("--ref-frames-per-step", type=int, default=0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
