"""Profile target: a short C2 ICP run (100k-point cylinder).  Run under ncu."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcreg_b200 import Context, default_params
from dcreg_b200.scenes import make_cylinder, g2_initial_pose

n = int(os.environ.get("ICP_POINTS", 100_000))
iters = int(os.environ.get("ICP_ITERS", 6))
pts = make_cylinder(n, seed=42)
with Context(0) as ctx:
    ctx.set_target(pts, float(os.environ.get("ICP_CELL", "1.0")))
    ctx.set_source(pts)
    prm = default_params(max_iterations=iters, fixed_iterations=1, kappa_target=10.0, use_weight_derivative=int(os.environ.get("ICP_WD", "0")))
    for _ in range(2):
        t0 = time.perf_counter()
        res = ctx.icp_run(prm, g2_initial_pose(), want_log=False)
        print("icp", res.iterations, "iterations", (time.perf_counter() - t0) * 1e3, "ms")
