"""The reference's real Franka capture (assets/franka_offline_example.zip: 10 views 640x480, annotated masks, qpos, K),
carried as tests/golden/franka_offline.npz (tools/make_fixture_franka.py): DAE link meshes + URDF forward kinematics
+ renderer against real silhouettes, and the GPU solver against an oracle-driven solve of the same data."""
import os
import sys

import numpy as np
import pytest
import torch

from easyhec_b200.meshio import Mesh
from easyhec_b200.se3 import dof_to_matrix
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIX = os.path.join(ROOT, "tests", "golden", "franka_offline.npz")


def load_fixture():
    d = np.load(FIX)
    H, W = int(d["H"]), int(d["W"])
    meshes = [Mesh(d[n + "_v"], d[n + "_f"]) for n in d["names"]]
    masks = np.unpackbits(d["masks_packed"], axis=-1)[:, :, :W].astype(bool)
    return d, meshes, masks, H, W


def pose_err(A, B):
    A, B = np.asarray(A, np.float64), np.asarray(B, np.float64)
    dR = A[:3, :3].T @ B[:3, :3]
    return float(np.linalg.norm(A[:3, 3] - B[:3, 3])), float(np.degrees(np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1))))


def test_fixture_is_the_reference_capture():
    d, meshes, masks, H, W = load_fixture()
    assert (H, W) == (480, 640) and masks.shape == (10, 480, 640) and d["qpos"].shape == (10, 9)
    assert [len(m.faces) for m in meshes] == [20483, 12516, 12716, 14233, 14621, 18327, 21620, 12082, 7078]
    assert np.allclose(d["K"], [[386.32171631, 0, 331.3142395], [0, 385.38577271, 239.80825806], [0, 0, 1]])
    assert 0.10 < masks.mean() < 0.16
    assert d["link_poses"].shape == (10, 9, 4, 4)
    # panda_link0 is the root: identity; panda_link1 sits 0.333 m above it, rotated about z by joint 1
    assert np.allclose(d["link_poses"][:, 0], np.eye(4), atol=1e-7)
    q1 = d["qpos"][:, 0]
    assert np.allclose(d["link_poses"][:, 1, 2, 3], 0.333, atol=1e-6)
    assert np.allclose(d["link_poses"][:, 1, 0, 0], np.cos(q1), atol=1e-6) and np.allclose(d["link_poses"][:, 1, 1, 0], np.sin(q1), atol=1e-6)


def test_real_silhouettes_overlap_the_rendered_robot():
    """DAE node transforms + URDF chain + projection conventions on real data: at the tuned pose the oracle's render
    overlaps the annotated masks (mean IoU 0.70; a wrong link frame or flipped axis drops it below 0.4)."""
    from util import mvp_of
    d, meshes, masks, H, W = load_fixture()
    packed = oracle.pack_links(meshes)
    mvp = np.stack([[mvp_of(d["K"], H, W, d["tuned_Tc_c2b"] @ d["link_poses"][b, l].astype(np.float64)) for l in range(9)]
                    for b in range(10)]).astype(np.float32)
    r = oracle.union_binary(packed, mvp, H, W)
    iou = np.array([(r[b] & masks[b]).sum() / (r[b] | masks[b]).sum() for b in range(10)])
    assert iou.mean() > 0.65 and iou.min() > 0.5, iou


def test_oracle_driven_solve_is_reproducible():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from make_fixture_franka import oracle_solve
    d, meshes, masks, H, W = load_fixture()
    traj, losses = oracle_solve(meshes, d["link_poses"], d["K"], masks, d["start_Tc_c2b"], 2, [0, 1, 2])
    it = list(d["traj_iters"])
    for k in (0, 1, 2):
        assert np.allclose(traj[k], d["traj_dof"][it.index(k)], rtol=0, atol=1e-6)
    assert abs(losses[0] - float(d["loss_values"][list(d["loss_iters"]).index(0)])) < 1e-3 * losses[0]


@pytest.mark.gpu
def test_gpu_solver_tracks_the_oracle_driven_solve_on_real_data():
    """PoseSolver (CUDA graph, fused kernels, device Adam) against torch.optim.Adam driven by the CPU oracle, same real
    masks, same start: the pose after 100 iterations agrees within the contract's 1 mm / 0.1 deg."""
    from easyhec_b200.solver import PoseSolver
    d, meshes, masks, H, W = load_fixture()
    s = PoseSolver(meshes, d["link_poses"], d["K"], masks, d["start_Tc_c2b"], H, W)
    n = int(d["traj_iters"].max())
    s.step(n)
    hist = s.history_ops().cpu().numpy()
    assert hist.shape == (n, 6)
    worst = (0.0, 0.0)
    for k, want in zip(d["traj_iters"], d["traj_dof"]):
        got = hist[k] if k < n else s.dof.detach().cpu().numpy()
        e = pose_err(dof_to_matrix(torch.as_tensor(got)).numpy(), dof_to_matrix(torch.as_tensor(want)).numpy())
        worst = (max(worst[0], e[0]), max(worst[1], e[1]))
        if k <= 2:
            assert e[0] < 2e-5 and e[1] < 2e-3, (k, e)     # the first steps are the same arithmetic
    assert worst[0] < 1e-3 and worst[1] < 0.1, worst
    # and it is an optimisation: the loss at the end is below the loss at the start
    assert float(s.loss) < float(d["loss_values"][0])
