"""K1 time vs slot count (slope = streaming rate, intercept = launch + final-reduction overhead)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcreg_b200 import Context
from dcreg_b200.scenes import make_corridor

nmax = int(os.environ.get("K1_SLOTS", 20_000_000))
scene = make_corridor(nmax, seed=44, noise=0.002)
T = np.eye(4); T[:3, 3] = [0.004, 0.003, -0.002]
with Context(0) as ctx:
    ctx.set_target(scene, 0.05)
    for n in (nmax, nmax // 2, nmax // 4, nmax // 10, nmax // 20, nmax // 100):
        ctx.set_source(scene[:n])
        ctx.find_planes(T, 0.05, want_planes=False)
        ctx.freeze_planes_f32()
        ctx.time_reduce(False, T, False, 3, False)
        t = ctx.time_reduce(False, T, False, 20, False)
        print(f"n={n:>9d}  {t*1e3:8.2f} us   {n*32/t/1e9:8.1f} GB/s")
