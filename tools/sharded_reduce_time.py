"""Sharded K1 (10 M slots over the ranks) in both exchange modes: CUDA-event time per launch incl. the sum over ranks."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from dcreg_b200 import Context
from dcreg_b200.parallel import init_sharded, shard_range
from dcreg_b200.scenes import make_corridor
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 10_000_000
scene = make_corridor(n, seed=44, noise=0.002)
T = np.eye(4); T[:3, 3] = [0.004, 0.003, -0.002]
lo, hi = shard_range(n, rank, world)
for mode in (2, 1):
    if mode == 1: os.environ["DCREG_NO_PEER"] = "1"
    else: os.environ.pop("DCREG_NO_PEER", None)
    ctx = Context(local)
    ctx.set_target(scene, 0.05); ctx.set_source(scene[lo:hi])
    ctx.find_planes(T, 0.05, want_planes=False); ctx.freeze_planes_f32()
    init_sharded(ctx, dist, n, device=dev)
    ctx.time_reduce(False, T, False, 5, False)
    ts = []
    for _ in range(5):
        dist.barrier(device_ids=[local]); torch.cuda.synchronize()
        t = torch.tensor([ctx.time_reduce(False, T, False, 50, False)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX); ts.append(float(t.item()) * 1e3)
    if rank == 0:
        print(f"N={world} mode {ctx.comm_mode}: K1 over {hi - lo} slots per rank + sum over ranks: " + " ".join(f"{x:.2f}" for x in ts) + " us per launch")
    ctx.close()
# the same shard WITHOUT any exchange (plain context): what the sum over ranks costs on top
ctx = Context(local)
ctx.set_target(scene, 0.05); ctx.set_source(scene[lo:hi])
ctx.find_planes(T, 0.05, want_planes=False); ctx.freeze_planes_f32()
ctx.time_reduce(False, T, False, 5, False)
ts = []
for _ in range(5):
    dist.barrier(device_ids=[local]); torch.cuda.synchronize()
    t = torch.tensor([ctx.time_reduce(False, T, False, 50, False)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX); ts.append(float(t.item()) * 1e3)
if rank == 0:
    print(f"N={world} no exchange: K1 over {hi - lo} slots per rank alone: " + " ".join(f"{x:.2f}" for x in ts) + " us per launch")
ctx.close()
dist.destroy_process_group()
