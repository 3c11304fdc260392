"""Run BASELINE.json's configs 2-5 on one B200 and print one JSON line each (kept under profiles/).

  cfg2  xArm7 links 1-7, 10 views 640x480, RBSolver pose optimisation, 200 Adam iterations (lr 3e-3, wd 5e-4):
        iterations/s of the CUDA-graph PoseSolver and the converged pose error against the generating pose
  cfg3  the real Franka visual meshes (link0-7 + hand, 133,676 triangles: tests/golden/franka_offline.npz), 20 views
        1280x720 with qpos ~ U(joint limits) (tests/golden/franka_cfg3.npz), fused render + loss + backward: frames/s
  cfg4  space exploration: 256 candidate joint configurations x 4 camera poses, xArm7 base + links 1-7
        (41,096 triangles), 1920x1080, binary render + variance score: candidates/s on ONE GPU (the path shards
        over ranks with one all-gather of the scores)
  cfg5  resolution (256^2 .. 2048^2) x views (1 .. 128) sweep, real Franka meshes, fwd+bwd: frames/s, algorithmic GB/s
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from easyhec_b200._lib import Context  # noqa: E402
from easyhec_b200.scenes import (FRANKA_K, chain_fk, franka_cfg3_scene, load_xarm7, make_scene, onscreen_fraction,  # noqa: E402
                                  perturb_pose, scaled_K)
from easyhec_b200.solver import PoseSolver  # noqa: E402
from util import scene_mvps  # noqa: E402


def timed(fn, n, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def timed_graph(fn, n, warm=3):
    """The same, with fn captured once into a CUDA graph and replayed (what the solver does): the host's launch rate -- a
    Python call per step -- does not bound small configurations."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g, stream=st):
            fn()
    torch.cuda.current_stream().wait_stream(st)
    torch.cuda.synchronize()
    return timed(g.replay, n, warm=warm)


def cfg2():
    B, H, W = 10, 480, 640
    sc = make_scene(B, H, W, links="xarm7", seed=0)
    ctx = Context("cuda:0")
    ids = [ctx.register_mesh(m.vertices, m.faces) for m in sc["meshes"]]
    ref = ctx.render_binary_batch(ids, torch.from_numpy(scene_mvps(sc, H, W)).cuda(), H, W)
    init = perturb_pose(sc["Tc_c2b"], np.random.RandomState(0), 0.03, 3.0)
    s = PoseSolver(sc["meshes"], sc["link_poses"], sc["K"], ref, init, H, W, ctx=ctx)
    e0 = s.pose_error(sc["Tc_c2b"])
    s.step(1)
    loss0 = float(s.loss)
    t0 = time.perf_counter()
    s.step(1)                      # captures the CUDA graph of one iteration, replays it once
    torch.cuda.synchronize()
    t_capture = time.perf_counter() - t0
    t0 = time.perf_counter()
    s.step(198)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    e1 = s.pose_error(sc["Tc_c2b"])
    out = {"config": "cfg2 xArm7 10 views 640x480, 200 Adam iterations", "iters_per_s": 198 / dt, "ms_per_iter": 1e3 * dt / 198,
           "graph_capture_ms": 1e3 * t_capture,
           "loss_first": loss0, "loss_last": float(s.loss), "init_err_mm_deg": [1e3 * e0[0], e0[1]],
           "final_err_mm_deg": [1e3 * e1[0], e1[1]]}
    s.step(800)
    e2 = s.pose_error(sc["Tc_c2b"])
    out["err_after_1000_iters_mm_deg"] = [1e3 * e2[0], e2[1]]
    out["loss_after_1000"] = float(s.loss)
    return out


def cfg3():
    B, H, W = 20, 720, 1280
    sc = franka_cfg3_scene(H, W, B)
    ctx = Context("cuda:0")
    ids = [ctx.register_mesh(m.vertices, m.faces) for m in sc["meshes"]]
    mvp_gt = torch.from_numpy(scene_mvps(sc, H, W)).cuda()
    mvp = torch.from_numpy(scene_mvps(sc, H, W, perturb_pose(sc["Tc_c2b"], np.random.RandomState(1), 0.03, 3.0))).cuda()
    ref_f = ctx.render_binary_batch(ids, mvp_gt, H, W).float()
    ref = ctx.register_ref(ref_f.to(torch.uint8))                       # registered once, like a solve does
    L = len(ids)
    out_t = (torch.empty((B, H, W), device="cuda"), torch.empty(B, dtype=torch.float64, device="cuda"),
             torch.empty((B, L, 4, 4), dtype=torch.float64, device="cuda"))
    ms = timed_graph(lambda: ctx.render_views_fused(ids, mvp, ref, H, W, backward=True, out=out_t), 100)
    flags, _ = ctx.status()
    V = sum(len(m.vertices) for m in sc["meshes"]); F = sum(len(m.faces) for m in sc["meshes"])
    alg = (8 * H * W + 40 * V + 24 * F) * B
    return {"config": "cfg3 real Franka visual meshes (%d triangles, 9 links), 20 views 1280x720, fwd+bwd" % F,
            "frames_per_s": B / (ms * 1e-3), "ms_per_step": ms, "flags": flags, "coverage": float(ref_f.mean()),
            "onscreen_frac": onscreen_fraction(sc, H, W), "algorithmic_GBps": alg / (ms * 1e-3) / 1e9}


def cfg4():
    Q, C, H, W = 256, 4, 1080, 1920
    fx = load_xarm7()
    rng = np.random.RandomState(0)
    lim = fx["joint_limits"]
    q = rng.uniform(np.maximum(lim[:, 0], -np.pi) * 0.6, np.minimum(lim[:, 1], np.pi) * 0.6, size=(Q, len(lim)))
    poses = np.stack([chain_fk(fx["joint_origin"], fx["joint_axis"], qq) for qq in q]).astype(np.float32)   # (Q,8,4,4)
    sc = make_scene(1, H, W, links="xarm7_all", seed=0, K_base=FRANKA_K)
    cams = [perturb_pose(sc["Tc_c2b"], np.random.RandomState(1 + c), 0.05, 5.0) for c in range(C)]
    from easyhec_b200.projection import K_to_projection, opencv2gl
    P = (K_to_projection(torch.from_numpy(scaled_K(H, W, FRANKA_K)), H, W) @ opencv2gl()).numpy()
    mvp = np.einsum("ij,cjk,qlkm->qclim", P, np.stack(cams).astype(np.float32), poses).astype(np.float32)   # (Q,C,L,4,4)
    ctx = Context("cuda:0")
    ids = [ctx.register_mesh(m.vertices, m.faces) for m in fx["meshes"]]
    mvp_d = torch.from_numpy(np.ascontiguousarray(mvp)).cuda()
    ms = timed(lambda: ctx.explore_scores(ids, mvp_d, H, W), 5, warm=1)
    score = ctx.explore_scores(ids, mvp_d, H, W).cpu().numpy()
    return {"config": "cfg4 space exploration: 256 qpos x 4 cameras, xArm7 base+1-7 (41,096 tri), 1920x1080, 1 GPU",
            "candidates_per_s": Q / (ms * 1e-3), "renders_per_s": Q * C / (ms * 1e-3), "ms_total": ms,
            "best_candidate": int(score.argmax()), "score_min_max": [float(score.min()), float(score.max())]}


def cfg5(resolutions=(256, 512, 1024, 2048), views=(1, 8, 32, 128)):
    """Resolution / batch sweep, Franka-sized robot, fused forward + backward: frames/s and algorithmic GB/s
    (8 H W + 40 V + 24 F bytes per frame) on one GPU.  The scratch pools start from their expected
    size at the large points; a raised overflow flag grows them and the step is run again (counted in `grows`)."""
    out = []
    for res in resolutions:
        for B in views:
            H = W = res
            sc = franka_cfg3_scene(H, W, B)
            sc["K"] = scaled_K(H, W)
            ctx = Context("cuda:0")
            ids = [ctx.register_mesh(m.vertices, m.faces) for m in sc["meshes"]]
            L = len(ids)
            V = sum(len(m.vertices) for m in sc["meshes"]); F = sum(len(m.faces) for m in sc["meshes"])
            mvp_gt = torch.from_numpy(scene_mvps(sc, H, W)).cuda()
            mvp = torch.from_numpy(scene_mvps(sc, H, W, perturb_pose(sc["Tc_c2b"], np.random.RandomState(1), 0.03, 3.0))).cuda()
            ref = ctx.register_ref(ctx.render_binary_batch(ids, mvp_gt, H, W))   # registered once, like a solve does
            out_t = (torch.empty((B, H, W), device="cuda"), torch.empty(B, dtype=torch.float64, device="cuda"),
                     torch.empty((B, L, 4, 4), dtype=torch.float64, device="cuda"))
            grows = 0
            while True:
                ctx.render_views_fused(ids, mvp, ref, H, W, backward=True, out=out_t)
                flags, _ = ctx.status()
                if not flags & 1:
                    break
                ctx.grow_scratch(); grows += 1
                assert grows < 8
            n = max(5, min(200, int(2000 / B)))
            ms = timed_graph(lambda: ctx.render_views_fused(ids, mvp, ref, H, W, backward=True, out=out_t), n)
            flags, _ = ctx.status()
            alg = (8 * H * W + 40 * V + 24 * F) * B
            out.append({"res": res, "views": B, "ms_per_step": round(ms, 4), "frames_per_s": round(B / (ms * 1e-3), 1),
                        "algorithmic_GBps": round(alg / (ms * 1e-3) / 1e9, 1), "flags": flags, "grows": grows})
            ctx.close()
            del ref, out_t, mvp, mvp_gt
            torch.cuda.empty_cache()
    return {"config": "cfg5 sweep: real Franka visual meshes (133,676 triangles, 9 links), fwd+bwd, 1 GPU", "points": out}


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg2", "cfg3", "cfg4", "cfg5"]
    for w in which:
        print(json.dumps(globals()[w]()), flush=True)
