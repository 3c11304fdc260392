"""Pose chain, drop-in renderer and RBSolver mirror on the GPU, against torch autograd / the oracle."""
import numpy as np
import pytest
import torch

from oracle import oracle
from easyhec_b200.scenes import SAMPLE_POSE, make_scene, perturb_pose, scaled_K
from util import mvp_of, rel_err, scene_mvps, to_dev

pytestmark = pytest.mark.gpu


def _scene(B=3, H=120, W=160, seed=11, links="xarm7"):
    sc = make_scene(B, H, W, links=links, seed=seed)
    packed = oracle.pack_links(sc["meshes"])
    ref = oracle.union_binary(packed, scene_mvps(sc, H, W), H, W)
    init = perturb_pose(sc["Tc_c2b"], np.random.RandomState(seed + 1), 0.02, 2.0)
    return sc, packed, ref, init


def test_pose_compose_and_backward_match_autograd(gpu_ctx):
    from easyhec_b200.rb_solver import compose_link_mvp
    from easyhec_b200.se3 import dof_to_matrix, matrix_to_dof
    H, W = 120, 160
    sc, _, _, init = _scene()
    dof = matrix_to_dof(torch.tensor(init, dtype=torch.float32)).cuda().requires_grad_(True)
    K = to_dev(sc["K"])
    lp = to_dev(sc["link_poses"])
    mvp_t = compose_link_mvp(K, H, W, dof_to_matrix(dof), lp)
    mvp_k = gpu_ctx.pose_compose(dof.detach().contiguous(), K, lp, H, W)
    assert rel_err(mvp_k.cpu().numpy(), mvp_t.detach().cpu().numpy()) < 2e-6
    g = torch.randn(mvp_t.shape, dtype=torch.float64, device="cuda")
    loss_b = torch.rand(lp.shape[0], dtype=torch.float64, device="cuda")
    (mvp_t.double() * g).sum().backward()
    out7 = gpu_ctx.pose_backward(dof.detach().contiguous(), K, lp, g.contiguous(), loss_b, H, W)
    assert rel_err(out7[:6].cpu().numpy(), dof.grad.cpu().numpy()) < 1e-4
    assert abs(out7[6].item() - loss_b.mean().item()) < 1e-6


def test_adam_step_matches_torch(gpu_ctx):
    p = torch.tensor([0.1, -0.2, 0.7, 0.3, -1.0, 0.5], device="cuda")
    ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=3e-3, weight_decay=5e-4)
    state = torch.zeros(13, device="cuda")
    hist = torch.zeros(8, 6, device="cuda")
    rng = torch.Generator(device="cpu").manual_seed(0)
    for it in range(5):
        g = torch.randn(7, generator=rng).cuda() * 10
        ref.grad = g[:6].clone()
        before = p.clone()
        opt.step()
        gpu_ctx.adam_step(p, g.contiguous(), state, 3e-3, weight_decay=5e-4, hist=hist)
        assert torch.allclose(p, ref.detach(), rtol=1e-6, atol=1e-7)
        assert torch.equal(hist[it], before)
    assert state[12].item() == 5


def test_pose_backward_adam_in_one_launch_equals_the_two_launches(gpu_ctx):
    """ehb_pose_backward_adam = ehb_pose_backward + ehb_adam_step_compose, bit for bit (same block arithmetic)."""
    g = torch.Generator(device="cpu").manual_seed(3)
    B, L, H, W = 5, 7, 120, 160
    K = torch.tensor([[150.0, 0, 80.0], [0, 151.0, 60.0], [0, 0, 1]], device="cuda")
    lp = torch.eye(4).repeat(B, L, 1, 1) + 0.1 * torch.randn(B, L, 4, 4, generator=g)
    lp[..., 3, :] = torch.tensor([0.0, 0, 0, 1]); lp = lp.cuda().contiguous()
    res = {}
    for fused in (False, True):
        dof = torch.tensor([0.1, -0.2, 0.7, 0.3, -1.0, 0.5], device="cuda")
        state = torch.zeros(13, device="cuda"); hist = torch.zeros(8, 6, device="cuda")
        mvp = torch.zeros(B, L, 4, 4, device="cuda")
        gg = torch.Generator(device="cpu").manual_seed(4)
        outs = []
        for it in range(3):
            g_mvp = torch.randn(B, L, 4, 4, generator=gg, dtype=torch.float64).cuda()
            loss = torch.rand(B, generator=gg, dtype=torch.float64).cuda()
            if fused:
                o7 = gpu_ctx.pose_backward_adam(dof, K, lp, g_mvp, loss, H, W, state, 3e-3, weight_decay=5e-4, hist=hist, mvp_next=mvp)
            else:
                o7 = gpu_ctx.pose_backward(dof, K, lp, g_mvp, loss, H, W)
                gpu_ctx.adam_step(dof, o7, state, 3e-3, weight_decay=5e-4, hist=hist, compose=(K, lp, H, W, mvp))
            outs.append(o7.clone())
        res[fused] = (torch.stack(outs), dof.clone(), state.clone(), hist.clone(), mvp.clone())
    for a, b in zip(res[False], res[True]):
        assert torch.equal(a, b)
    assert res[True][2][12].item() == 3 and res[True][4].abs().sum() > 0
    # Adam on a copy of the parameters (the bench's scratch): the pose parameters stay
    dof = torch.tensor([0.1, -0.2, 0.7, 0.3, -1.0, 0.5], device="cuda"); scratch = dof.clone(); keep = dof.clone()
    gpu_ctx.pose_backward_adam(dof, K, lp, g_mvp, loss, H, W, torch.zeros(13, device="cuda"), 3e-3, adam_dof=scratch)
    assert torch.equal(dof, keep) and not torch.equal(scratch, keep)


def test_rbsolver_fused_vs_loop_vs_oracle():
    from easyhec_b200.rb_solver import RBSolver, compose_link_mvp
    from easyhec_b200.se3 import dof_to_matrix
    H, W = 120, 160
    sc, packed, ref, init = _scene()
    dps = {"mask": torch.from_numpy(ref).cuda().float(), "link_poses": to_dev(sc["link_poses"]),
           "K": to_dev(sc["K"])[None], "Tc_c2b": torch.tensor(sc["Tc_c2b"], dtype=torch.float32)[None].cuda(),
           "global_step": 0}
    out = {}
    for fused in (True, False):
        m = RBSolver(meshes=sc["meshes"], init_Tc_c2b=init, H=H, W=W, fused=fused)
        o, ld = m(dps)
        ld["mask_loss"].backward()
        out[fused] = (ld["mask_loss"].item(), m.dof.grad.cpu().numpy().copy(), o["rendered_masks"].detach().cpu().numpy(),
                      o["metrics"]["err_trans"].item())
        mvp = compose_link_mvp(dps["K"][0], H, W, dof_to_matrix(m.dof.detach()), dps["link_poses"]).cpu().numpy()
    want = oracle.render_views(packed, mvp, ref.astype(np.float32), H, W)
    assert np.array_equal(out[True][2], want["masks"])
    # the drop-in operator, link by link: its mvp is associated differently (Tc @ lp per link), so the blend
    # weights may differ in the last ulp; coverage is identical
    assert np.abs(out[False][2] - want["masks"]).max() < 1e-5
    assert np.array_equal(out[False][2] > 0.5, want["masks"] > 0.5)
    assert abs(out[True][0] - want["loss"]) < 1e-6 * want["loss"]
    assert abs(out[False][0] - want["loss"]) < 1e-5 * want["loss"]
    assert rel_err(out[True][1], out[False][1]) < 1e-4            # north_star: gradients within 1e-4 relative
    assert 0.5 < out[True][3] < 6.0                               # ~2 cm perturbation reported in cm


def test_dropin_renderer_matches_oracle_and_autograd():
    from easyhec_b200.renderer import NVDiffrastRenderer
    H, W = 96, 128
    sc = make_scene(1, H, W, links="xarm7", seed=5)
    li = 1     # a link that is in view for this seed (asserted below through want.max())
    m = sc["meshes"][li]
    K = to_dev(sc["K"])
    pose = torch.tensor(sc["Tc_c2b"] @ sc["link_poses"][0, li].astype(np.float64), dtype=torch.float32).cuda()
    pose.requires_grad_(True)
    verts, faces = to_dev(m.vertices), to_dev(m.faces)
    r = NVDiffrastRenderer([H, W])
    mask = r.render_mask(verts, faces, K, pose)
    mvp = mvp_of(sc["K"], H, W, pose.detach().cpu().numpy())
    want, st = oracle.render_mask(m.vertices, m.faces, mvp, H, W, anti_aliasing=True, save=True)
    assert mask.dtype == torch.float32 and mask.shape == (H, W)
    assert np.abs(mask.detach().cpu().numpy() - want).max() < 1e-6 and want.max() > 0.5
    dy = torch.rand(H, W, device="cuda")
    (mask * dy).sum().backward()
    _, g_mvp = oracle.render_mask_bwd(m.vertices, m.faces, mvp, H, W, st, dy.cpu().numpy())
    P = mvp_of(sc["K"], H, W, np.eye(4)).astype(np.float64)
    assert rel_err(pose.grad.cpu().numpy(), P.T @ g_mvp) < 1e-4
    b = r.render_mask(verts, faces, K, pose.detach(), anti_aliasing=False)
    assert b.dtype == torch.bool
    assert np.array_equal(b.cpu().numpy(), oracle.render_mask(m.vertices, m.faces, mvp, H, W, anti_aliasing=False))
    # batch_render_mask: vertices already in the camera frame
    vc = (verts @ pose.detach()[:3, :3].T + pose.detach()[:3, 3]).contiguous()
    b2 = r.batch_render_mask(vc, faces, K, anti_aliasing=False)
    want2 = oracle.render_mask(vc.cpu().numpy(), m.faces, mvp_of(sc["K"], H, W, np.eye(4)), H, W, anti_aliasing=False)
    assert np.array_equal(b2.cpu().numpy(), want2) and want2.sum() > 0     # the oracle on the same camera-frame vertices
    # the projection is kept per intrinsics tensor: a second call gives the same mask, an in-place change of K a new one
    assert torch.equal(r.render_mask(verts, faces, K, pose.detach()), mask.detach())
    K2 = K.clone(); r.render_mask(verts, faces, K2, pose.detach()); K2[0, 2] += 7.0
    shifted = r.render_mask(verts, faces, K2, pose.detach())
    assert not torch.equal(shifted, mask.detach())
    # the oracle on the matrix the operator composed (a 4x4 product on the device may round differently from the host's in the last
    # ulp, which an antialiased edge magnifies); that matrix against the host composition separately
    mvp3 = (r._proj_flip(K2) @ pose.detach()).cpu().numpy()
    assert np.allclose(mvp3, mvp_of(K2.cpu().numpy(), H, W, pose.detach().cpu().numpy()), rtol=1e-5, atol=1e-6)
    want3 = oracle.render_mask(m.vertices, m.faces, mvp3, H, W, anti_aliasing=True)
    d3 = np.abs(shifted.cpu().numpy() - want3).max()
    assert d3 < 1e-6, d3
    with pytest.raises(RuntimeError):
        r.render_mask(verts, faces.long(), K, pose)        # int64 faces are rejected like nvdiffrast does


def test_pose_solver_converges_and_graph_equals_eager():
    from easyhec_b200.solver import PoseSolver
    H, W = 120, 160
    sc, _, ref, init = _scene(B=4, seed=21)
    kw = dict(meshes=sc["meshes"], link_poses=sc["link_poses"], K=sc["K"], masks_ref=ref, init_Tc_c2b=init, H=H, W=W)
    a = PoseSolver(use_graph=True, **kw)
    b = PoseSolver(use_graph=False, **kw)
    e0 = a.pose_error(sc["Tc_c2b"])
    a.step(30); b.step(30)
    assert torch.allclose(a.dof, b.dof, rtol=1e-5, atol=1e-6)
    a.step(170)
    e1 = a.pose_error(sc["Tc_c2b"])
    assert e1[0] < 0.35 * e0[0] and e1[0] < 0.006, (e0, e1)
    assert a.history_ops().shape == (200, 6)
    assert float(a.loss) >= 0


@pytest.mark.gpu
def test_steps_in_flight_on_slots_match_the_single_stream_path(gpu_ctx):
    """ehb_step_begin: two rounds of four steps over the four slots (device matrices, masks out, pose chain + Adam behind the pass) give
    the oracle's masks and the same 7 floats / Adam update as the single-stream calls; host matrices + host outputs too."""
    from easyhec_b200.se3 import matrix_to_dof
    B, H, W = 3, 240, 320
    sc = make_scene(B, H, W, links="xarm7", seed=9)
    packed = oracle.pack_links(sc["meshes"])
    ref = oracle.union_binary(packed, scene_mvps(sc, H, W), H, W)
    ids = [gpu_ctx.register_mesh(m.vertices, m.faces) for m in sc["meshes"]]
    h = gpu_ctx.register_ref(ref)
    K = to_dev(sc["K"]); lp = to_dev(sc["link_poses"])
    sets = []
    for k in range(4):
        Tc = perturb_pose(sc["Tc_c2b"], np.random.RandomState(50 + k), 0.02, 2.0)
        mvp = scene_mvps(sc, H, W, Tc)
        want = oracle.render_views(packed, mvp, ref.astype(np.float32), H, W)
        dof = matrix_to_dof(torch.tensor(Tc, dtype=torch.float32)).cuda().contiguous()
        _, l1, g1 = gpu_ctx.render_views_fused(ids, to_dev(mvp), h, H, W, backward=True, want_masks=False)
        g7 = gpu_ctx.pose_backward(dof, K, lp, g1, l1, H, W, grad_scale=1.0, loss_scale=1.0 / B)
        d2 = torch.zeros(6, device="cuda"); st2 = torch.zeros(13, device="cuda")
        gpu_ctx.adam_step(d2, g7, st2, 3e-3, weight_decay=5e-4)
        sets.append(dict(mvp=mvp, want=want, dof=dof, g7=g7.clone(), adam=d2))
    masks = [torch.full((B, H, W), -1.0, device="cuda") for _ in range(4)]
    o7 = [torch.zeros(7, device="cuda") for _ in range(4)]
    ad = [torch.zeros(6, device="cuda") for _ in range(4)]
    st = [torch.zeros(13, device="cuda") for _ in range(4)]
    mvp_dev = [to_dev(s["mvp"]) for s in sets]
    for rnd in range(2):                                     # the second round reuses every slot and overwrites the outputs
        for j in range(4):
            ad[j].zero_(); st[j].zero_(); masks[j].fill_(-1.0)
        gpu_ctx.slots_fork()                                 # the slot streams wait for the zeroing above
        for j in range(4):
            gpu_ctx.step_begin(j, ids, h, H, W, mvp_dev[j], masks=masks[j], dof=sets[j]["dof"], K=K, link_poses=lp, out7=o7[j],
                               adam_dof=ad[j], adam_state=st[j], lr=3e-3, weight_decay=5e-4)
        gpu_ctx.slots_join()
        torch.cuda.synchronize()
        for j in range(4):
            gpu_ctx.solver_step_end(j)
            assert np.array_equal(masks[j].cpu().numpy(), sets[j]["want"]["masks"])
            assert torch.allclose(o7[j], sets[j]["g7"], rtol=1e-5, atol=1e-7)
            assert torch.allclose(ad[j], sets[j]["adam"], rtol=1e-5, atol=1e-8)
    # host matrices in, host results out (what the end-to-end leg of the bench uses), on two slots at once
    loss_h = [torch.empty(B, dtype=torch.float64).pin_memory() for _ in range(2)]
    g_h = [torch.empty((B, len(ids), 4, 4), dtype=torch.float64).pin_memory() for _ in range(2)]
    o7_h = [torch.empty(7, dtype=torch.float32).pin_memory() for _ in range(2)]
    mvp_h = [torch.from_numpy(sets[j]["mvp"]).pin_memory() for j in range(2)]
    for j in range(2):
        gpu_ctx.step_begin(j, ids, h, H, W, mvp_h[j], loss_host=loss_h[j], g_mvp_host=g_h[j], dof=sets[j]["dof"], K=K, link_poses=lp,
                           out7_host=o7_h[j])
    for j in range(2):
        gpu_ctx.solver_step_end(j)
        assert np.allclose(loss_h[j].numpy(), sets[j]["want"]["loss_per_view"], rtol=1e-12, atol=0)
        assert rel_err(g_h[j].numpy(), sets[j]["want"]["g_mvp"]) < 1e-9
        assert torch.allclose(o7_h[j], sets[j]["g7"].cpu(), rtol=1e-5, atol=1e-7)
