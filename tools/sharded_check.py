"""2+ rank check (torchrun): sharded ICP (point blocks + ncclAllReduce of 32 doubles) == single-GPU ICP."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from dcreg_b200 import Context, default_params
from dcreg_b200.parallel import init_sharded
from dcreg_b200.scenes import make_cylinder, g2_initial_pose

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pts = make_cylinder(60_000, seed=42)
T0 = g2_initial_pose()
prm = default_params(max_iterations=30, fixed_iterations=1, kappa_target=10.0)   # long enough to reach the record-reusing mode
ctx = Context(local)
ctx.set_target(pts, 1.0)
# single-GPU reference on every rank
ctx.set_source(pts)
ref = ctx.icp_run(prm, T0)
# sharded
lo, hi = init_sharded(ctx, dist, len(pts), device=torch.device("cuda", local))
ctx.set_source(pts[lo:hi])
ctx.set_global_source_count(len(pts))
res = ctx.icp_run(prm, T0)
dT = np.abs(res.T - ref.T).max()
ok = res.iterations == ref.iterations and dT < 1e-9 and all(
    a.n_effective == b.n_effective and abs(a.fitness - b.fitness) < 1e-12 for a, b in zip(res.logs, ref.logs))
t = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("max |T_sharded - T_single| =", dT, "SHARDED_OK" if t.item() == 1.0 else "SHARDED_MISMATCH")
ctx.close()
dist.destroy_process_group()
sys.exit(0 if t.item() == 1.0 else 1)
