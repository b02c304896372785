"""Where one loop iteration spends its time: per-block phase stamps of the iteration kernel + the solve step (C2 scene)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcreg_b200 import Context, default_params
from dcreg_b200.scenes import make_cylinder, g2_initial_pose

pts = make_cylinder(int(os.environ.get("ICP_POINTS", 100_000)), seed=42)
T0 = g2_initial_pose()
prm = default_params(kappa_target=10.0)
names = ["start", "cert", "search", "fitlist", "fits", "gram", "reduced", "sums", "solved"]
with Context(0) as ctx:
    ctx.set_target(pts, 1.0)
    ctx.set_source(pts)
    for iters in (1, 3, 8, 14, 25, 40):
        blocks, solve = ctx.iteration_timeline(prm, T0, iters)
        t0 = blocks[:, 0].min()
        last = int(solve[15])
        print(f"iteration {iters - 1}: {len(blocks)} blocks; block starts span {(blocks[:, 0].max() - t0) / 1e3:.1f} us")
        for k in range(1, 6):
            col = blocks[:, k] - t0
            print(f"   {names[k]:8s} done: min {col.min() / 1e3:6.1f}  median {np.median(col) / 1e3:6.1f}  max {col.max() / 1e3:6.1f} us after the first block start")
        d = np.diff(blocks[:, :6], axis=1)
        print("   per-phase duration, median / max over blocks (us):", "  ".join(f"{names[k + 1]} {np.median(d[:, k]) / 1e3:.1f}/{d[:, k].max() / 1e3:.1f}" for k in range(5)))
        ws = blocks[:, 9:14].sum(axis=0).astype(float)
        if ws[3] > 0:
            print(f"   warp search (warp 0 of every block, {int(ws[3])} searches): set-up {ws[0] / ws[3]:.0f}  scan {ws[1] / ws[3]:.0f}  select {ws[2] / ws[3]:.0f} cycles, {ws[4] / ws[3]:.0f} candidates per search")
        lb = blocks[last]
        print(f"   last block {last}: gram done at {(lb[5] - t0) / 1e3:.1f}, reduced {(lb[6] - t0) / 1e3:.1f}, sums {(lb[7] - t0) / 1e3:.1f}, solved {(lb[8] - t0) / 1e3:.1f} us")
        print("   solve step (us): " + "  ".join(f"{n} {(solve[k + 1] - solve[k]) / 1e3:.2f}" for k, n in enumerate(["inverses", "schur+jacobi", "precond", "pcg", "update"])),
              f" total {(solve[5] - solve[0]) / 1e3:.2f}")
