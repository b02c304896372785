#!/usr/bin/env python
"""bench.py - ICP iterations/s (whole-job) and Mpoints/s of the J^T J / J^T r reduction on B200.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W [--impl reference]`
prints ONE JSON line on rank 0.  A "step" is one full registration of the C2 workload (synthetic
100k-point cylinder pair, 50 fixed ICP iterations: correspondences + K1 + K2 every iteration).

  value      ICP iterations/s, source + target resident in HBM when the timed region starts
  e2e        same metric through the public C-ABI call with HOST buffers: every step uploads the
             scan from pinned host memory and reads back pose + per-iteration log
  roofline   K1 (fused residual/weight/Jacobian/27-sum reduction) at C4 size (10M slots, frozen
             float4 planes, 32 B/slot), CUDA-event timed on the launching stream, vs MEASURED_PEAKS
  cpu_baseline  the CPU oracle (port of the reference loop) timed on this box's host cores on a
             bounded sample of the same workload
N > 1: replicas (one independent scan pair per GPU, no data-path collective, "weak"); the sharded
10M-slot reduction with its 32-double ncclAllReduce is reported under "sharded".
`--impl reference` times the reference's own CPU algorithm (oracle port; the reference binary cannot
be built here: Eigen/PCL/yaml-cpp absent and its "Ours" stage is a stub) on the same config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C2_POINTS = 100_000
C2_ITERS = 50
C4_SLOTS = 10_000_000
C4_RADIUS = 0.05
ALG_BYTES_PER_SLOT = 32          # float4 point + float4 plane
K1_NCU_TRAFFIC_BYTES = 320.06e6 + 3.93e6   # dram read + write of one 10 M-slot K1 launch (ncu --set full, profiles/k1_r1_final_ncu_summary.txt)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def c2_params(default_params):
    # "Ours" = [SCHUR_CONDITION_NUMBER, PRECONDITIONED_CG] (icp_pk01.yaml:106), kappa_th = kappa_tg = 10 (icp.yaml),
    # USE_WEIGHT_DERIVATIVE = false (the released default, icp_test_runner.cpp:1691; with the derivative term the
    # Gauss-Newton iteration DIVERGES on this 100k scene - checked with the CPU oracle - so it is not a sane
    # benchmark workload), init = the published cylinder perturbation, 50 fixed iterations (BASELINE.json configs[1])
    return default_params(search_radius=1.0, max_iterations=C2_ITERS, fixed_iterations=1, kappa_target=10.0,
                          cond_thresh=10.0, use_weight_derivative=0, detection="SCHUR_CONDITION_NUMBER",
                          handling="PRECONDITIONED_CG")


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# ------------------------------------------------------------------------------------------------
_CPU_SCENE = {}


def cpu_icp_sample(n_points, iters, seed, thread_mode=1):
    """Time `iters` full ICP iterations of the C/OpenMP oracle (the port of the reference loop) on the C2 scene,
    kd-tree build excluded as in the reference's own timing (icp_test_runner.cpp:408-461).
    thread_mode 1 = all host threads (correspondences + 27-sum reduction), 0 = reference-faithful (8 threads on
    correspondences only, serial Jacobian build and A^T A).  Returns (seconds, threads, kind)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dcreg_oracle_c as oc
    from dcreg_b200.scenes import make_cylinder, g2_initial_pose
    key = (n_points, seed)
    if key not in _CPU_SCENE:
        pts = make_cylinder(n_points, seed=seed)
        _CPU_SCENE[key] = (pts, oc.Scene(pts, pts))
    pts, sc = _CPU_SCENE[key]
    prm = oc.make_params(max_iterations=iters, fixed_iterations=True, kappa_target=10.0, use_weight_derivative=False,
                         thread_mode=thread_mode)
    t0 = time.perf_counter()
    st, conv, n_it, T, _ = sc.icp_run(prm, g2_initial_pose(), want_log=False)
    dt = time.perf_counter() - t0
    assert st == 0 and n_it == iters
    return dt, (8 if thread_mode == 0 else oc.max_threads()), "port"


def run_reference(args, rank, world):
    """The reference's own CPU algorithm (oracle port) on the host cores; rank 0 only."""
    if rank != 0:
        return
    sample_iters = 10
    for _ in range(max(0, min(args.warmup, 2))):
        cpu_icp_sample(C2_POINTS, 2, 42)
    times = []
    cores = 1
    for _ in range(max(1, args.steps)):
        sec, cores, kind = cpu_icp_sample(C2_POINTS, sample_iters, 42)
        times.append(sec)
    tot = float(np.sum(times))
    value = sample_iters * len(times) / tot
    sec0, cores0, _ = cpu_icp_sample(C2_POINTS, 5, 42, thread_mode=0)
    line = {
        "impl": "reference", "metric": "icp_iterations_per_s", "value": value, "unit": "ICP iterations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C2 synthetic cylinder pair {C2_POINTS} pts, method Ours (Schur detection + PCG), "
                               "search_radius 1.0; CPU port of the reference loop (reference binary not buildable here)",
                   "sample": f"{sample_iters} ICP iterations per step (the GPU arm runs {C2_ITERS} per step)"},
        "cpu_baseline": {"value": value, "unit": "ICP iterations/s", "cores": cores, "kind": "port",
                         "sample": f"{sample_iters} iterations x {len(times)} steps of the C2 workload, all host threads",
                         "reference_faithful_8_threads": {"value": 5 / sec0, "cores": cores0,
                                                          "note": "omp num_threads(8) on correspondences only, serial J build and A^T A (icp_test_runner.cpp:1714,1863-1915)"}},
        "e2e": {"value": value, "unit": "ICP iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from dcreg_b200 import Context, default_params
    from dcreg_b200.scenes import make_cylinder, make_corridor, g2_initial_pose

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - dcreg_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    # ---------------- C2: full ICP iterations/s (replicas at N > 1) ----------------
    pts = make_cylinder(C2_POINTS, seed=42 + rank)
    pinned = torch.from_numpy(pts).pin_memory()
    pts_pinned = pinned.numpy()
    T0 = g2_initial_pose()
    prm = c2_params(default_params)
    ctx.set_target(pts, 1.0)                 # cell = radius (measured fastest; finer grids are exact too but slower): index build: setup, outside the reference's timed region too
    ctx.set_source(pts_pinned)
    for _ in range(max(args.warmup, 3)):
        res = ctx.icp_run(prm, T0, want_log=False)
    assert res.iterations == C2_ITERS

    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        res = ctx.icp_run(prm, T0, want_log=False)       # inputs resident in HBM
    e1.record(stream)
    e1.synchronize()
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launch_count - l0
    value = world * args.steps * C2_ITERS / (dev_ms * 1e-3)

    # how much correspondence work the loop reused in one step (untimed extra run, counters on)
    ctx.iteration_counters(True)
    ctx.icp_run(prm, T0, want_log=False)
    searched, fitted = ctx.iteration_counters(False)

    # e2e: host buffers in, pose + log out, every step
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e2.record(stream)
    for _ in range(args.steps):
        ctx.set_source(pts_pinned)                        # H2D of this step's scan (pinned)
        res = ctx.icp_run(prm, T0, want_log=True)         # D2H of pose + per-iteration records
    e3.record(stream)
    e3.synchronize()
    wall = time.perf_counter() - w0
    barrier()
    e2e_ms = max_over_ranks(max(e2.elapsed_time(e3), wall * 1e3))
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = world * args.steps * C2_ITERS / (e2e_ms * 1e-3)
    from dcreg_b200.api import IterLog
    import ctypes
    h2d = int(pts.shape[0] * 3 * 4 + 16 * 8)
    d2h = int(ctypes.sizeof(IterLog) * C2_ITERS + 472)

    # ---------------- C4: K1 reduction roofline (sharded at N > 1) ----------------
    from dcreg_b200.parallel import init_sharded, shard_range
    n_total = C4_SLOTS
    scene = make_corridor(n_total, seed=44, noise=0.002)
    lo, hi = shard_range(n_total, rank, world)
    Tc = np.eye(4); Tc[:3, 3] = [0.004, 0.003, -0.002]
    ctx.set_target(scene, C4_RADIUS)
    ctx.set_source(scene[lo:hi])
    ctx.find_planes(Tc, C4_RADIUS, want_planes=False)     # correspondences once; planes stay on the device
    ctx.freeze_planes_f32()
    if world > 1:
        init_sharded(ctx, dist, n_total, device=dev)      # NCCL comm inside the C library; id via torch.distributed
    else:
        ctx.set_global_source_count(n_total)
    out27, stats = ctx.reduce_device(False, Tc, False)    # warm-up + sanity
    for wd in (False, True):
        for f64 in (False, True):
            ctx.time_reduce(f64, Tc, wd, 3, False)
    barrier()
    reps = 20
    k1_ms = max_over_ranks(ctx.time_reduce(False, Tc, False, reps, False))     # inputs (320 MB) > L2 (126 MB)
    k1_ms_f64 = max_over_ranks(ctx.time_reduce(True, Tc, False, reps, False))
    k1_ms_wd = max_over_ranks(ctx.time_reduce(False, Tc, True, reps, False))
    barrier()
    peak, peak_src = measured_peak_gbs()
    n_local = hi - lo
    achieved = ALG_BYTES_PER_SLOT * n_local / (k1_ms * 1e-3) / 1e9             # per GPU
    mpts = n_total / (k1_ms * 1e-3) / 1e6                                       # whole job

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec, cores, kind = cpu_icp_sample(C2_POINTS, 3, 42)                 # calibrate
        it_all = int(min(200, max(10, 12.0 / (sec / 3))))                    # ~12 s of CPU work
        sec, cores, kind = cpu_icp_sample(C2_POINTS, it_all, 42)
        it_ref = int(min(100, max(5, it_all // 3)))
        sec0, cores0, _ = cpu_icp_sample(C2_POINTS, it_ref, 42, thread_mode=0)
        cpu = {"value": it_all / sec, "unit": "ICP iterations/s", "cores": cores, "kind": kind,
               "sample": f"{it_all} ICP iterations of the C2 workload, all host threads (kd-tree build excluded, as in the reference)",
               "reference_faithful_8_threads": {"value": it_ref / sec0, "cores": cores0, "sample": f"{it_ref} iterations",
                                                "note": "omp num_threads(8) on correspondences only, serial J build and A^T A"}}

    if rank == 0:
        line = {
            "metric": "icp_iterations_per_s", "value": value, "unit": "ICP iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C2 synthetic cylinder pair {C2_POINTS} pts x {C2_ITERS} fixed ICP iterations per step, "
                                   "method Ours (Schur detection + PCG), weight derivative off (released default), device grid correspondences",
                       "parallelism": "replicas (one scan pair per GPU)" if world > 1 else "1 GPU",
                       "l2": "K1 roofline inputs 320 MB > 126 MB L2; no flush needed",
                       "roofline_workload": f"C4 synthetic corridor {C4_SLOTS} slots, frozen float4 planes"},
            "e2e": {"value": e2e_value, "unit": "ICP iterations/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "loop": {"slot_iterations_per_step": C2_POINTS * C2_ITERS, "searched": int(searched), "plane_fits": int(fitted),
                     "note": "every iteration recomputes every correspondence; a slot whose 7 stored neighbours provably still "
                             "contain its 5 nearest (gap certificate) skips the cell search, a slot whose 5 neighbours are the same "
                             "ordered list reuses its plane - results identical to searching and fitting every time (tests/test_gpu_parity.py)"},
            "roofline": {"kernel": "k1s::reduce_stream_kernel<float4, wd=false> (K1)", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": K1_NCU_TRAFFIC_BYTES * n_local / C4_SLOTS, "traffic_source": "ncu --set full dram__bytes_read+write per 10 M-slot launch, profiles/k1_r1_final_ncu_summary.txt",
                         "peak_source": peak_src,
                         "ms_per_launch": k1_ms, "slots_per_launch": n_local, "bytes_per_slot": ALG_BYTES_PER_SLOT},
            "reduction": {"mpoints_per_s": mpts, "slots_total": n_total, "ms": k1_ms,
                          "weight_derivative_variant_ms": k1_ms_wd,
                          "weight_derivative_variant_gbs": ALG_BYTES_PER_SLOT * n_local / (k1_ms_wd * 1e-3) / 1e9,
                          "f64_plane_variant_ms": k1_ms_f64,
                          "f64_plane_variant_gbs": 48 * n_local / (k1_ms_f64 * 1e-3) / 1e9,
                          "n_effective": int(stats[1]),
                          "collective": "ncclAllReduce 32 doubles per launch (inside the timed region)" if world > 1 else None,
                          "sharding": f"{world} contiguous point blocks of {n_local} slots" if world > 1 else None},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
