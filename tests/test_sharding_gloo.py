"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: the shard split, the unique-id broadcast, and the
property the sharded path relies on - per-block normal-equation sums add up to the whole (checked with the oracle)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_tiles_exactly():
    from dcreg_b200.parallel import shard_range
    for n in (0, 1, 7, 100_000, 10_000_001):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import dcreg_oracle as o
        from dcreg_b200.parallel import broadcast_unique_id, shard_range
        from dcreg_b200.scenes import g2_initial_pose
        uid = broadcast_unique_id(lambda: bytes(range(128)), dist)
        assert uid == bytes(range(128))
        pts = o.read_pcd_xyz(os.path.join(ROOT, "tests", "golden", "cylinder_7562.pcd"))
        T = g2_initial_pose()
        corr = o.find_correspondences(pts, pts, o.build_tree(pts), T[:3, :3], T[:3, 3], 1.0, True)
        plane = np.concatenate([corr.n, corr.d[:, None]], axis=1); plane[~corr.valid] = 0
        src4 = np.concatenate([pts, np.zeros((len(pts), 1), np.float32)], axis=1)
        lo, hi = shard_range(len(pts), rank, world)
        part, st = o.reduce_normal_equations(src4[lo:hi], plane[lo:hi], T[:3, :3], T[:3, 3], True)
        t = torch.from_numpy(np.concatenate([part, st]))
        dist.all_reduce(t)                                   # what ncclAllReduce(sum, double) does on the GPUs
        whole, stw = o.reduce_normal_equations(src4, plane, T[:3, :3], T[:3, 3], True)
        ref = np.concatenate([whole, stw])
        ok = np.abs(t.numpy() - ref).max() <= 1e-12 * np.abs(ref).max()
        # every rank then solves redundantly on identical sums -> identical update
        H, g = o.unpack27(t.numpy()[:27])
        prm = o.Params(kappa_target=10.0, use_weight_derivative=True)
        dx = o.solve_degenerate_system(H, g, prm, o.analyze_degeneracy(H, prm))
        gathered = [torch.zeros(6, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(dx))
        same = all(torch.equal(gathered[0], x) for x in gathered)
        q.put((rank, bool(ok), bool(same)))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_sums_and_id_broadcast():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1] and all(r[1] and r[2] for r in res)


@pytest.mark.gpu
def test_sharded_icp_matches_single_gpu():
    """Needs >= 2 GPUs: 2 ranks, point-block sharding + the 32-double all-reduce == the single-GPU run."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29731",
                          os.path.join(ROOT, "tools", "sharded_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "SHARDED_OK" in out.stdout
