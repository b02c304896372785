"""Per-kernel timing of the fused ICP loop on the C2 workload (CUDA events, not under a profiler).
Also prints how many source slots searched / refitted in each of the first iterations."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcreg_b200 import Context, default_params
from dcreg_b200.scenes import make_cylinder, g2_initial_pose

n = int(os.environ.get("ICP_POINTS", 100_000))
reps = int(os.environ.get("ICP_REPS", 50))
pts = make_cylinder(n, seed=42)
with Context(0) as ctx:
    ctx.set_target(pts, float(os.environ.get("ICP_CELL", "1.0")))
    ctx.set_source(pts)
    prm = default_params(kappa_target=10.0, use_weight_derivative=int(os.environ.get("ICP_WD", "0")))
    T0 = g2_initial_pose()
    res = ctx.icp_run(default_params(max_iterations=reps, fixed_iterations=1, kappa_target=10.0), T0, want_log=False)
    for name, T in (("initial pose", T0), ("converged pose", res.T)):
        it = ctx.time_iteration(prm, T, 0, reps)
        print(f"{name}: iteration kernel alone {it*1e3:7.2f} us")
    both = ctx.time_iteration(prm, T0, 1, reps)
    print(f"{reps} real iterations: {both*1e3:7.2f} us per iteration (iteration kernel + solve kernel)")
    for k in (1, 2, 3, 4, 6, 8, 12, 16, 24, 50):
        t = ctx.time_iteration(prm, T0, 1, k)
        print(f"  first {k:2d} iterations: {t*1e3:7.2f} us per iteration")
    ctx.iteration_counters(True)
    for k in range(1, 17):
        ctx.icp_run(default_params(max_iterations=k, fixed_iterations=1, kappa_target=10.0), T0, want_log=False)
        s, f = ctx.iteration_counters(True)
        if k > 1:
            print(f"  iteration {k-1:2d}: searched {s - prev[0]:6d}  refitted {f - prev[1]:6d} of {n}")
        prev = (s, f)
    ctx.iteration_counters(False)
    # late regime: start from the pose after 30 iterations
    res30 = ctx.icp_run(default_params(max_iterations=30, fixed_iterations=1, kappa_target=10.0), T0, want_log=False)
    ctx.iteration_counters(True)
    for k in (1, 5, 20):
        t = ctx.time_iteration(prm, res30.T, 1, k)
        s, f = ctx.iteration_counters(True)
        print(f"  from the pose after 30 iterations, {k:2d} iterations: {t*1e3:7.2f} us per iteration; searched {s} refitted {f} (incl. 2 warm-up launches)")
    t = ctx.time_iteration(prm, res30.T, 0, 20)
    print(f"  pose after 30 iterations, iteration kernel alone (fixed pose): {t*1e3:7.2f} us")
    ctx.iteration_counters(False)
