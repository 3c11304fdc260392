"""Camera model of the render_mask operator (host side).

Reference: easyhec/utils/nvdiffrast_utils.py:5-11 `K_to_projection`, :14-18 `transform_pos`;
easyhec/structures/nvdiffrast_renderer.py:18-22,33-37 (OpenCV->GL flip and ``proj @ pose``).

The CUDA kernels and the oracle both take the composed 4x4 ``mvp = proj @ flip @ object_pose``
(row-major, fp32) and apply it per vertex in one fixed operation order, so every float that
decides a pixel is identical on both sides; this module only builds ``mvp``.
"""
import numpy as np
import torch

__all__ = ["K_to_projection", "opencv2gl", "compose_mvp", "transform_pos", "NEAR", "FAR"]

NEAR, FAR = 0.001, 10.0


def K_to_projection(K, H: int, W: int, n: float = NEAR, f: float = FAR) -> torch.Tensor:
    """Pinhole intrinsics -> GL projection, fp32 (4,4) on K's device.

    Entries that depend on K are computed in fp32 step by step (the reference multiplies 0-d
    fp32 tensors), the depth row in Python floats then rounded to fp32 -- same as the reference.
    No ``.item()``: the matrix is assembled with tensor ops, so there is no device sync.
    """
    K = torch.as_tensor(K, dtype=torch.float32)
    fu, fv, cu, cv = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    z = torch.zeros((), dtype=torch.float32, device=K.device)
    a = torch.tensor(-(f + n) / (f - n), dtype=torch.float32, device=K.device)
    b = torch.tensor(-2 * f * n / (f - n), dtype=torch.float32, device=K.device)
    m1 = torch.tensor(-1.0, dtype=torch.float32, device=K.device)
    rows = [2 * fu / W, z, -2 * cu / W + 1, z,
            z, 2 * fv / H, 2 * cv / H - 1, z,
            z, z, a, b,
            z, z, m1, z]
    return torch.stack(rows).reshape(4, 4)


def opencv2gl(device=None) -> torch.Tensor:
    """diag(1,-1,-1,1): camera looks down -Z with +Y up (nvdiffrast_renderer.py:18-22)."""
    return torch.diag(torch.tensor([1.0, -1.0, -1.0, 1.0], device=device))


def compose_mvp(K, H: int, W: int, object_pose: torch.Tensor) -> torch.Tensor:
    """``proj @ (opencv2gl @ object_pose)``; object_pose (...,4,4) -> mvp (...,4,4), differentiable."""
    proj = K_to_projection(K, H, W).to(object_pose.device)
    return proj @ (opencv2gl(object_pose.device) @ object_pose)


def transform_pos(mtx: torch.Tensor, pos: torch.Tensor) -> torch.Tensor:
    """(V,3) object-space points -> (1,V,4) clip space, ``[p,1] @ mtx^T`` (nvdiffrast_utils.py:14-18)."""
    if isinstance(mtx, np.ndarray):
        mtx = torch.from_numpy(mtx).to(pos.device)
    ones = torch.ones([pos.shape[0], 1], dtype=pos.dtype, device=pos.device)
    return torch.matmul(torch.cat([pos, ones], dim=1), mtx.t())[None, ...]
