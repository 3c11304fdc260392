"""world_size-2 gloo test of the multi-GPU host logic: views are sharded round-robin, every rank reduces its own
views to (d loss/d mvp, loss), and ONE 7-float all-reduce yields the global-mean gradient (trainer/base.py:349).
The per-rank compute is stood in for by the CPU oracle here; the arithmetic of the exchange is what is tested."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from easyhec_b200.solver import shard_views


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle
    from easyhec_b200.scenes import make_scene, perturb_pose
    from util import scene_mvps
    B, H, W = 5, 60, 80          # 5 views over 2 ranks: uneven shards (3 + 2)
    sc = make_scene(B, H, W, links="xarm7", seed=9)
    packed = oracle.pack_links(sc["meshes"])
    ref = oracle.union_binary(packed, scene_mvps(sc, H, W), H, W).astype(np.float32)
    mvp = scene_mvps(sc, H, W, perturb_pose(sc["Tc_c2b"], np.random.RandomState(3), 0.02, 2.0))
    mine = shard_views(B, rank, world)
    local = oracle.render_views(packed, mvp[mine], ref[mine], H, W)          # gradient of the LOCAL mean
    # what PoseSolver does: rescale by B_local / B_global, loss by 1 / B_global, then all-reduce(sum)
    g = torch.from_numpy(local["g_mvp"].sum(axis=(0, 1)).reshape(-1) * (len(mine) / B))
    l = torch.tensor([local["loss_per_view"].sum() / B])
    buf = torch.cat([g, l])
    dist.all_reduce(buf)
    if rank == 0:
        full = oracle.render_views(packed, mvp, ref, H, W)
        q.put((buf.numpy(), full["g_mvp"].sum(axis=(0, 1)).reshape(-1), full["loss"]))
    dist.destroy_process_group()


def test_shard_views_partition():
    for n, w in [(10, 2), (5, 2), (7, 8), (256, 8)]:
        parts = [shard_views(n, r, w) for r in range(w)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_two_rank_allreduce_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, want_g, want_l = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the per-pixel gradient is scaled by 1/B in fp32 (as torch does for the mean): local and global 1/B round differently
    assert np.allclose(got[:-1], want_g, rtol=2e-6, atol=1e-9)
    assert np.isclose(got[-1], want_l, rtol=1e-12)
