"""CUDA vs oracle on the configurations that are BENCHMARKED (BASELINE.json configs, bench.py's workloads), at their full
sizes: the bench scene itself (every ring slot), the real Franka meshes of config 3, the exploration scorer of config 4
end to end, and a 2048^2 point of the config-5 sweep.  Masks bit-exact, loss 1e-12, gradients 1e-9 relative (the
contract is 1e-4)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import oracle
from easyhec_b200.scenes import FRANKA_K, franka_cfg3_scene, load_xarm7, make_scene, perturb_pose, scaled_K
from util import rel_err, scene_mvps, to_dev

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    return b


def _check_fused(ctx, ids, packed, mvp, ref, H, W, u8=False):
    want = oracle.render_views(packed, mvp, ref, H, W)
    r = to_dev(ref.astype(np.uint8)) if u8 else to_dev(ref)
    masks, loss, g_mvp = ctx.render_views_fused(ids, to_dev(mvp), r, H, W, backward=True)
    flags, nclip = ctx.status()
    assert flags & 1 == 0
    got = masks.cpu().numpy()
    assert np.array_equal(got, want["masks"]), "%d pixels differ" % int((got != want["masks"]).sum())
    assert np.allclose(loss.cpu().numpy(), want["loss_per_view"], rtol=1e-12, atol=0)
    assert want["loss"] > 0 and np.abs(want["g_mvp"]).max() > 0
    assert rel_err(g_mvp.cpu().numpy(), want["g_mvp"]) < 1e-9
    assert nclip == want["n_need_clip"]
    # the same against REGISTERED reference masks (bit-packed once, per-tile counts): identical results
    h = ctx.register_ref(ref.astype(np.uint8) if u8 else ref)
    masks2, loss2, g2 = ctx.render_views_fused(ids, to_dev(mvp), h, H, W, backward=True)
    _, loss3, g3 = ctx.render_views_fused(ids, to_dev(mvp), h, H, W, backward=True, want_masks=False)
    flags, _ = ctx.status()
    assert flags & 1 == 0
    assert np.array_equal(masks2.cpu().numpy(), want["masks"])
    for l_, g_ in ((loss2, g2), (loss3, g3)):
        assert np.allclose(l_.cpu().numpy(), want["loss_per_view"], rtol=1e-12, atol=0)
        assert rel_err(g_.cpu().numpy(), want["g_mvp"]) < 1e-9
    h.release()
    return want


@pytest.mark.parametrize("workload", ["headline", "inview"])
def test_bench_scene_every_ring_slot(gpu_ctx, workload):
    """bench.py's own inputs (build_sets: 10 views x 1280x720, xArm7 links 1-7), all four ring slots."""
    b = _bench()
    wl = b.WORKLOADS[workload]
    H, W = wl["H"], wl["W"]
    sets = b.build_sets(wl, 0, wl["ring"])
    packed = oracle.pack_links(sets[0]["scene"]["meshes"])
    ids = [gpu_ctx.register_mesh(m.vertices, m.faces) for m in sets[0]["scene"]["meshes"]]
    for k, s in enumerate(sets):
        ref = oracle.union_binary(packed, s["mvp_gt"], H, W)
        # the bench builds its reference masks with the binary GPU path: same bits
        got_ref = gpu_ctx.render_binary_batch(ids, to_dev(s["mvp_gt"]), H, W).cpu().numpy().astype(bool)
        assert np.array_equal(got_ref, ref), "slot %d" % k
        _check_fused(gpu_ctx, ids, packed, s["mvp"], ref.astype(np.float32), H, W, u8=bool(k & 1))
    for i in ids:
        gpu_ctx.release_mesh(i)


def test_cfg3_real_franka_meshes_20_views_1280x720(gpu_ctx):
    """BASELINE.json config 3 on the real Franka visual meshes (133,676 triangles, 9 links)."""
    H, W = 720, 1280
    sc = franka_cfg3_scene(H, W)
    assert sum(len(m.faces) for m in sc["meshes"]) == 133676 and sc["link_poses"].shape == (20, 9, 4, 4)
    packed = oracle.pack_links(sc["meshes"])
    ref = oracle.union_binary(packed, scene_mvps(sc, H, W), H, W)
    assert 0.03 < ref.mean() < 0.2
    mvp = scene_mvps(sc, H, W, perturb_pose(sc["Tc_c2b"], np.random.RandomState(1), 0.03, 3.0))
    ids = [gpu_ctx.register_mesh(m.vertices, m.faces) for m in sc["meshes"]]
    _check_fused(gpu_ctx, ids, packed, mvp, ref.astype(np.float32), H, W)
    for i in ids:
        gpu_ctx.release_mesh(i)


def test_cfg4_score_candidates_end_to_end_1920x1080(gpu_ctx, tmp_path):
    """BASELINE.json config 4 through the product's own entry point (explore.score_candidates: URDF forward kinematics ->
    mvp of every (candidate, camera, link) -> binary render -> variance score), against the oracle's union render +
    variance on the same matrices.  A subset of the 256 candidates keeps the CPU side at a few seconds."""
    from easyhec_b200.explore import candidate_mvps, score_candidates
    from util import xarm_urdf
    H, W, Q, C = 1080, 1920, 12, 4
    fx = load_xarm7()
    kin = xarm_urdf(tmp_path, fx)
    rng = np.random.RandomState(0)
    lim = fx["joint_limits"]
    q = rng.uniform(np.maximum(lim[:, 0], -np.pi) * 0.6, np.minimum(lim[:, 1], np.pi) * 0.6, size=(Q, len(lim)))
    sc = make_scene(1, H, W, links="xarm7_all", seed=0, K_base=FRANKA_K)
    cams = np.stack([perturb_pose(sc["Tc_c2b"], np.random.RandomState(1 + c), 0.05, 5.0) for c in range(C)])
    K = scaled_K(H, W, FRANKA_K)
    links = list(range(8))
    ids = [gpu_ctx.register_mesh(m.vertices, m.faces) for m in fx["meshes"]]
    valid = np.ones(Q, bool); valid[3] = False
    packed = oracle.pack_links(fx["meshes"])
    # (a) matrices composed on the host (torch): the device render + variance against the oracle on the same matrices
    got = score_candidates(gpu_ctx, ids, kin, links, q, cams, K, H, W, valid=valid, device_fk=False).cpu().numpy()
    flags, _ = gpu_ctx.status()
    assert flags & 1 == 0
    mvp = candidate_mvps(kin, links, q, cams, K, H, W).numpy()
    masks = oracle.union_binary(packed, mvp.reshape(Q * C, len(links), 4, 4), H, W)
    want = oracle.variance_scores(masks.reshape(Q, C, H, W))
    want[3] = 0.0
    assert want.max() > 0
    assert np.allclose(got, want, rtol=1e-14, atol=0)
    # (b) forward kinematics + composition on the device (ehb_explore_fk_mvp): its matrices agree with the host's to fp32
    # rounding, and the scores are exact for the matrices it produced
    robot = gpu_ctx.register_robot(kin)
    mvp_d = gpu_ctx.explore_fk_mvp(robot, to_dev(q), cams, K, H, W, links)
    assert mvp_d.shape == (Q, C, len(links), 4, 4)
    assert np.abs(mvp_d.cpu().numpy() - mvp).max() <= 2e-6 * np.abs(mvp).max()
    got_d = score_candidates(gpu_ctx, ids, kin, links, q, cams, K, H, W, valid=valid, robot=robot).cpu().numpy()
    masks_d = oracle.union_binary(packed, mvp_d.cpu().numpy().reshape(Q * C, len(links), 4, 4), H, W)
    want_d = oracle.variance_scores(masks_d.reshape(Q, C, H, W))
    want_d[3] = 0.0
    assert np.allclose(got_d, want_d, rtol=1e-14, atol=0)
    assert np.abs(got_d - want).max() < 0.02 * want.max()       # and both describe the same candidates
    for i in ids:
        gpu_ctx.release_mesh(i)


def test_cfg5_point_2048x2048(gpu_ctx):
    """One point of the config-5 sweep: 2048 x 2048, 2 views, xArm7 links (the real meshes), fwd + bwd."""
    H = W = 2048
    sc = make_scene(2, H, W, links="xarm7", seed=4)
    packed = oracle.pack_links(sc["meshes"])
    ref = oracle.union_binary(packed, scene_mvps(sc, H, W), H, W)
    mvp = scene_mvps(sc, H, W, perturb_pose(sc["Tc_c2b"], np.random.RandomState(2), 0.02, 2.0))
    ids = [gpu_ctx.register_mesh(m.vertices, m.faces) for m in sc["meshes"]]
    _check_fused(gpu_ctx, ids, packed, mvp, ref.astype(np.float32), H, W)
    for i in ids:
        gpu_ctx.release_mesh(i)
