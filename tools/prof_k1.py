"""Profile target: a few K1 launches at C4 size (10M slots).  Run under ncu:  ncu ... python tools/prof_k1.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcreg_b200 import Context
from dcreg_b200.scenes import make_corridor

n = int(os.environ.get("K1_SLOTS", 10_000_000))
reps = int(os.environ.get("K1_REPS", 3))
scene = make_corridor(n, seed=44, noise=0.002)
T = np.eye(4); T[:3, 3] = [0.004, 0.003, -0.002]
with Context(0) as ctx:
    ctx.set_target(scene, 0.05)
    ctx.set_source(scene)
    ctx.find_planes(T, 0.05, want_planes=False)
    ctx.freeze_planes_f32()
    for wd in (False, True):
        for f64 in (False, True):
            ctx.time_reduce(f64, T, wd, 2, False)          # one-time per-kernel setup outside the timed batch
    for wd in (False, True):
        print("f32 planes wd", wd, ctx.time_reduce(False, T, wd, reps, False), "ms")
        print("f64 planes wd", wd, ctx.time_reduce(True, T, wd, reps, False), "ms")
