"""Per-kernel timing of the fused ICP loop on the C2 workload (CUDA events, not under a profiler).
DCREG_IT_DEBUG=1/2/3 ablates the plane fit / the search / the grid reduction inside the iteration kernel."""
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
