"""50-iteration C2 run time for the current DCREG_COOP_MAX / DCREG_COHERENT_STEP settings (CUDA events via time_iteration)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dcreg_b200 import Context, default_params
from dcreg_b200.scenes import make_cylinder, g2_initial_pose
pts = make_cylinder(100_000, seed=42)
with Context(0) as ctx:
    ctx.set_target(pts, 1.0); ctx.set_source(pts)
    prm = default_params(kappa_target=10.0)
    T0 = g2_initial_pose()
    ctx.time_iteration(prm, T0, 1, 50)
    t50 = ctx.time_iteration(prm, T0, 1, 50)
    t20 = ctx.time_iteration(prm, T0, 1, 20)
    print(f"coop_max={os.environ.get('DCREG_COOP_MAX','-')} step={os.environ.get('DCREG_COHERENT_STEP','-')}: "
          f"50 its {t50*1e3:6.2f} us/it, first 20 {t20*1e3:6.2f}, last 30 {(50*t50-20*t20)/30*1e3:6.2f}")
