"""Host-side geometry against goldens produced by RUNNING the reference's own Python
(tools/make_golden_host_math.py -> tests/golden/host_math.npz)."""
import os

import numpy as np
import torch

from easyhec_b200.projection import K_to_projection, compose_mvp, transform_pos
from easyhec_b200.se3 import dof_to_matrix, matrix_to_dof, se3_exp_map, se3_log_map

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "host_math.npz"))


def test_K_to_projection_bit_exact():
    for K, (H, W), want in zip(G["K"], G["HW"], G["proj"]):
        got = K_to_projection(torch.from_numpy(K), int(H), int(W)).numpy()
        assert got.dtype == np.float32 and np.array_equal(got, want)


def test_se3_exp_map_matches_reference():
    got = se3_exp_map(torch.from_numpy(G["dof"])).numpy()
    assert np.allclose(got, G["exp"], rtol=0, atol=2e-7)


def test_se3_log_map_opencv_backend_matches_reference():
    got = se3_log_map(torch.from_numpy(G["exp"])).numpy()
    assert np.allclose(got, G["log"], rtol=0, atol=2e-6)
    # theta -> 0 rows: the eps clamp, not a Taylor branch, handles them
    assert np.allclose(got[0, 3:], 0, atol=1e-6)


def test_rbsolver_init_dof_matches_reference():
    dof = matrix_to_dof(G["init_Tc_c2b"]).numpy()
    assert np.allclose(dof, G["init_dof"], atol=2e-6)
    back = dof_to_matrix(torch.from_numpy(dof)).numpy()
    assert np.allclose(back, G["init_Tc_c2b"], atol=5e-6)


def test_render_mask_clip_chain_matches_reference():
    mvp = compose_mvp(torch.from_numpy(G["K"][0]), 720, 1280, torch.from_numpy(G["chain_pose"])).numpy()
    assert np.allclose(mvp, G["chain_mvp"], rtol=0, atol=1e-6)
    clip = transform_pos(torch.from_numpy(G["chain_mvp"]), torch.from_numpy(G["chain_verts"]))[0].numpy()
    assert np.allclose(clip, G["chain_clip"], rtol=1e-6, atol=1e-6)
    # the oracle / kernels apply mvp with a fixed fma order: same values to fp32 rounding
    from oracle import oracle
    clip2 = oracle.transform(G["chain_verts"], G["chain_mvp"])
    assert np.allclose(clip2, G["chain_clip"], rtol=2e-6, atol=2e-6)


def test_se3_exp_is_differentiable_and_round_trips():
    dof = torch.tensor([0.1, -0.2, 0.8, 0.3, -1.1, 0.4], requires_grad=True)
    T = dof_to_matrix(dof)
    assert torch.allclose(T[:3, :3] @ T[:3, :3].T, torch.eye(3), atol=1e-6)
    T.sum().backward()
    assert dof.grad is not None and torch.isfinite(dof.grad).all()
    assert torch.allclose(matrix_to_dof(T.detach()), dof.detach(), atol=1e-5)
