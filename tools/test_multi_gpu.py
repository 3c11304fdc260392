"""Multi-GPU checks (run with torchrun on a box with >= 2 GPUs; not part of the single-GPU pytest run):
  1. ehb_allreduce7 (NVLink peer mailboxes) == NCCL all-reduce, bit for bit across ranks, over many steps;
  2. PoseSolver with views sharded over the ranks reaches the same pose as a single-rank solve of all views.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/test_multi_gpu.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from easyhec_b200._lib import Context
    from easyhec_b200.scenes import make_scene, perturb_pose
    from easyhec_b200.solver import PoseSolver, shard_views
    from util import scene_mvps
    ctx = Context(dev)
    ctx.comm_connect()
    g = torch.Generator(device="cpu").manual_seed(rank)
    worst = 0.0
    for step in range(200):
        v = torch.randn(7, generator=g).to(dev)
        a = v.clone(); b = v.clone()
        ctx.allreduce7(a)
        dist.all_reduce(b)
        worst = max(worst, float((a - b).abs().max()))
        gathered = [torch.empty_like(a) for _ in range(world)]
        dist.all_gather(gathered, a)
        assert all(torch.equal(gathered[0], t) for t in gathered), "ranks disagree"
    assert worst < 1e-5, worst
    # the exchange fused into the pose chain: pose_backward(send) + adam(recv) leaves the NCCL sum of the ranks' out7 in g7
    Bq, Lq = 3, 2
    gen = torch.Generator(device="cpu").manual_seed(100 + rank)
    dofq = (torch.randn(6, generator=gen) * 0.3).to(dev)
    Kq = torch.tensor([[100.0, 0, 64], [0, 100.0, 48], [0, 0, 1]], device=dev)
    lpq = torch.eye(4).repeat(Bq, Lq, 1, 1).to(dev).contiguous()
    for step in range(20):
        gm = torch.randn(Bq, Lq, 4, 4, generator=gen, dtype=torch.float64).to(dev)
        lb = torch.rand(Bq, generator=gen, dtype=torch.float64).to(dev)
        ref7 = ctx.pose_backward(dofq, Kq, lpq, gm, lb, 96, 128)
        dist.all_reduce(ref7)
        g7 = ctx.pose_backward(dofq, Kq, lpq, gm, lb, 96, 128, send=True)
        ctx.adam_step(dofq.clone(), g7, torch.zeros(13, device=dev), 0.0, recv=True)
        assert torch.allclose(g7, ref7, rtol=1e-5, atol=1e-6), (g7, ref7)
        gathered = [torch.empty_like(g7) for _ in range(world)]
        dist.all_gather(gathered, g7)
        assert all(torch.equal(gathered[0], t) for t in gathered), "ranks disagree (fused exchange)"
    if rank == 0:
        print("fused exchange (pose_backward send + adam recv) == NCCL all-reduce, identical on every rank")
    # steps in flight on the context's slots: every slot exchanges on its own mailbox channel; 12 steps over 4 slots leave the
    # NCCL sum of the ranks' out7 in every slot's buffer, and the Adam parameters agree on every rank
    Hs, Ws, Bs = 120, 160, 2
    scs = make_scene(Bs, Hs, Ws, links="xarm7", seed=31 + rank)
    ids_s = [ctx.register_mesh(m.vertices, m.faces) for m in scs["meshes"]]
    ref_s = ctx.register_ref(ctx.render_binary_batch(ids_s, torch.from_numpy(scene_mvps(scs, Hs, Ws)).to(dev), Hs, Ws))
    from easyhec_b200.se3 import matrix_to_dof
    Tcs = perturb_pose(scs["Tc_c2b"], np.random.RandomState(32 + rank), 0.02, 2.0)
    mvp_s = torch.from_numpy(scene_mvps(scs, Hs, Ws, Tcs)).to(dev)
    dof_s = matrix_to_dof(torch.tensor(Tcs, dtype=torch.float32)).to(dev).contiguous()
    K_s = torch.from_numpy(scs["K"]).to(dev).contiguous()
    lp_s = torch.from_numpy(scs["link_poses"]).to(dev).contiguous()
    _, l_ref, g_ref = ctx.render_views_fused(ids_s, mvp_s, ref_s, Hs, Ws, backward=True, want_masks=False)
    want7 = ctx.pose_backward(dof_s, K_s, lp_s, g_ref, l_ref, Hs, Ws, grad_scale=1.0 / world, loss_scale=1.0 / (Bs * world))
    dist.all_reduce(want7)
    o7 = [torch.zeros(7, device=dev) for _ in range(4)]
    ad = [torch.zeros(6, device=dev) for _ in range(4)]
    st = [torch.zeros(13, device=dev) for _ in range(4)]
    ctx.slots_fork()
    for k in range(12):
        ctx.step_begin(k % 4, ids_s, ref_s, Hs, Ws, mvp_s, dof=dof_s, K=K_s, link_poses=lp_s, out7=o7[k % 4], adam_dof=ad[k % 4],
                       adam_state=st[k % 4], lr=1e-3, grad_scale=1.0 / world, loss_scale=1.0 / (Bs * world), exchange=True)
    ctx.slots_join()
    torch.cuda.synchronize()
    for k in range(4):
        ctx.solver_step_end(k)
        assert torch.allclose(o7[k], want7, rtol=1e-5, atol=1e-6), (k, o7[k], want7)
        gathered = [torch.empty_like(ad[k]) for _ in range(world)]
        dist.all_gather(gathered, ad[k])
        assert all(torch.equal(gathered[0], t) for t in gathered) and float(ad[k].abs().max()) > 0, "ranks disagree (slot exchange)"
    if rank == 0:
        print("12 steps in flight over 4 slots: every slot's exchange == NCCL all-reduce, Adam parameters identical on every rank")
    # latency of the two collectives (device time, 200 back-to-back calls)
    for name, fn in (("peer", lambda t: ctx.allreduce7(t)), ("nccl", lambda t: dist.all_reduce(t))):
        t = torch.zeros(7, device=dev)
        for _ in range(20):
            fn(t)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            fn(t)
        e1.record(); torch.cuda.synchronize()
        if rank == 0:
            print("%s all-reduce of 7 floats: %.2f us per call" % (name, 1e3 * e0.elapsed_time(e1) / 200))
    # sharded solve vs single-rank solve
    B, H, W = 6, 120, 160
    sc = make_scene(B, H, W, links="xarm7", seed=21)
    ids = [ctx.register_mesh(m.vertices, m.faces) for m in sc["meshes"]]
    ref = ctx.render_binary_batch(ids, torch.from_numpy(scene_mvps(sc, H, W)).to(dev), H, W)
    init = perturb_pose(sc["Tc_c2b"], np.random.RandomState(22), 0.02, 2.0)
    mine = shard_views(B, rank, world)
    s = PoseSolver(sc["meshes"], sc["link_poses"][mine], sc["K"], ref[mine], init, H, W, n_views_global=B, device=dev)
    s.step(100)
    full = PoseSolver(sc["meshes"], sc["link_poses"], sc["K"], ref, init, H, W, device=dev, group=None, peer_allreduce=False)
    full.world = 1
    full.step(100)
    d = float((s.dof - full.dof).abs().max())
    e = s.pose_error(sc["Tc_c2b"])
    if rank == 0:
        print("sharded vs single-rank dof max |diff| = %.2e; pose error %.3f mm / %.4f deg (peer all-reduce: %s)" %
              (d, 1e3 * e[0], e[1], s._peer))
    assert d < 2e-3, d
    # space exploration: candidates block-partitioned over the ranks + ONE all-gather == all candidates on one rank
    from easyhec_b200.explore import score_candidates
    from easyhec_b200.scenes import load_xarm7
    from util import xarm_urdf
    import pathlib
    import tempfile
    fx = load_xarm7()
    kin = xarm_urdf(pathlib.Path(tempfile.mkdtemp()), fx)
    Hx, Wx, Qn, Cn = 270, 480, 11, 3                      # 11 candidates over the ranks: uneven blocks
    qs = np.random.RandomState(5).uniform(-1.5, 1.5, size=(Qn, 7))
    cams = np.stack([perturb_pose(sc["Tc_c2b"], np.random.RandomState(40 + k), 0.05, 5.0) for k in range(Cn)])
    from easyhec_b200.scenes import scaled_K
    Kx = scaled_K(Hx, Wx)
    ids8 = [ctx.register_mesh(m.vertices, m.faces) for m in fx["meshes"]]
    sharded = score_candidates(ctx, ids8, kin, list(range(8)), qs, cams, Kx, Hx, Wx)
    was = dist.is_initialized
    dist.is_initialized = lambda: False                   # the same call as a single-rank job
    try:
        single = score_candidates(ctx, ids8, kin, list(range(8)), qs, cams, Kx, Hx, Wx)
    finally:
        dist.is_initialized = was
    assert sharded.shape == (Qn,) and torch.equal(sharded, single) and float(single.max()) > 0
    if rank == 0:
        print("exploration scores sharded over %d ranks + all-gather == single rank (%d candidates)" % (world, Qn))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("multi-GPU checks ok")


if __name__ == "__main__":
    main()
