"""SE(3) log/exp parametrisation of the camera pose (host side, PyTorch).

Restates the maps RBSolver uses for its 6-DoF ``dof`` parameter
(reference: easyhec/utils/pytorch3d_se3.py:12-41 `_so3_exp_map`, :46-130 `se3_exp_map`,
:218-258 `_se3_V_matrix/_get_se3_V_input`; easyhec/utils/utils_3d.py:308-335 `se3_log_map`
with ``backend='opencv'``).  Conventions kept from the reference:

* ``dof = [t(3) | w(3)]``; the 4x4 returned by :func:`se3_exp_map` is in the *row-vector*
  convention (translation in the last row), so callers transpose it
  (rb_solver.py:52 does ``.permute(0, 2, 1)``).
* ``theta = sqrt(clamp(|w|^2, eps))`` -- the clamp, not a Taylor branch, handles theta -> 0.
* ``hat(w) = [[0,-z,y],[z,0,-x],[-y,x,0]]`` (pytorch3d.transforms.so3.hat; not vendored).

Everything is differentiable torch code: this is plumbing, the renderer is the product.
"""
import math

import numpy as np
import torch

__all__ = ["hat", "se3_exp_map", "se3_log_map", "so3_log_rodrigues", "dof_to_matrix", "matrix_to_dof"]


def hat(w: torch.Tensor) -> torch.Tensor:
    """(N,3) -> (N,3,3) skew-symmetric matrices, h @ v == cross(w, v)."""
    if w.ndim != 2 or w.shape[1] != 3:
        raise ValueError("Input vectors have to be of shape (N, 3).")
    x, y, z = w.unbind(1)
    o = torch.zeros_like(x)
    return torch.stack([o, -z, y, z, o, -x, -y, x, o], dim=1).reshape(-1, 3, 3)


def _angles(w: torch.Tensor, eps: float) -> torch.Tensor:
    return torch.clamp((w * w).sum(1), eps).sqrt()


def _V_matrix(w_hat, w_hat2, theta):
    eye = torch.eye(3, dtype=w_hat.dtype, device=w_hat.device)[None]
    c1 = ((1 - torch.cos(theta)) / (theta ** 2))[:, None, None]
    c2 = ((theta - torch.sin(theta)) / (theta ** 3))[:, None, None]
    return eye + w_hat * c1 + w_hat2 * c2


def se3_exp_map(log_transform: torch.Tensor, eps: float = 1e-4) -> torch.Tensor:
    """(N,6) -> (N,4,4), row-vector convention: ``[[R^T, 0], [T, 1]]``."""
    if log_transform.ndim != 2 or log_transform.shape[1] != 6:
        raise ValueError("Expected input to be of shape (N, 6).")
    t = log_transform[:, :3]
    w = log_transform[:, 3:]
    theta = _angles(w, eps)
    inv = 1.0 / theta
    f1 = inv * theta.sin()
    f2 = inv * inv * (1.0 - theta.cos())
    K = hat(w)
    K2 = torch.bmm(K, K)
    R = f1[:, None, None] * K + f2[:, None, None] * K2 + torch.eye(3, dtype=w.dtype, device=w.device)[None]
    V = _V_matrix(K, K2, theta)
    T = torch.bmm(V, t[:, :, None])[:, :, 0]
    out = torch.zeros(log_transform.shape[0], 4, 4, dtype=log_transform.dtype, device=log_transform.device)
    out[:, :3, :3] = R
    out[:, :3, 3] = T
    out[:, 3, 3] = 1.0
    return out.permute(0, 2, 1)


def so3_log_rodrigues(R: np.ndarray) -> np.ndarray:
    """Rotation matrix -> rotation vector, the quantity cv2.Rodrigues(R)[0] returns.

    Uses OpenCV when importable (what the reference calls, utils_3d.py:320-322); otherwise a
    numpy restatement (axis from the skew part, angle from atan2) that agrees to ~1e-7.
    """
    R = np.asarray(R, dtype=np.float64)
    try:
        import cv2
        return cv2.Rodrigues(R)[0].reshape(3)
    except Exception:  # pragma: no cover - OpenCV is present in the image
        rx, ry, rz = R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]
        s = 0.5 * math.sqrt(rx * rx + ry * ry + rz * rz)
        c = 0.5 * (np.trace(R) - 1.0)
        th = math.atan2(s, c)
        if s < 1e-12:
            return np.zeros(3) if c > 0 else th * np.sqrt(np.maximum((np.diag(R) + 1) / 2, 0))
        return np.array([rx, ry, rz]) * (th / (2 * s))


def se3_log_map(transform: torch.Tensor, eps: float = 1e-4, test_acc: bool = True) -> torch.Tensor:
    """(N,4,4) row-vector SE(3) -> (N,6); the reference's ``backend='opencv'`` branch.

    ``w = -Rodrigues(transform[:3,:3])`` (the block is R^T, hence the sign),
    ``t = V(w)^-1 T`` with ``T = transform[3,:3]``.
    """
    if transform.ndim != 3 or transform.shape[1:] != (4, 4):
        raise ValueError("Input tensor shape has to be (N, 4, 4).")
    ws = []
    for m in transform:
        rv = -so3_log_rodrigues(m[:3, :3].detach().cpu().numpy())
        ws.append(torch.from_numpy(rv.reshape(-1)).to(transform.device).float())
    w = torch.stack(ws, 0)
    T = transform[:, 3, :3]
    K = hat(w)
    V = _V_matrix(K, torch.bmm(K, K), _angles(w, eps))
    t = torch.linalg.solve(V, T[:, :, None])[:, :, 0]
    dof = torch.cat((t, w), dim=1)
    if test_acc:
        err = (se3_exp_map(dof) - transform).abs().max()
        if err > 0.1:
            raise RuntimeError("se3_log_map round trip error %g > 0.1" % float(err))
    return dof


def dof_to_matrix(dof: torch.Tensor) -> torch.Tensor:
    """(6,) -> (4,4) column-vector pose ``Tc_c2b`` exactly as rb_solver.py:52 builds it."""
    return se3_exp_map(dof[None]).permute(0, 2, 1)[0]


def matrix_to_dof(Tc_c2b, eps: float = 1e-5) -> torch.Tensor:
    """(4,4) pose -> (6,) dof, as rb_solver.py:32-33 initialises the parameter."""
    T = torch.as_tensor(Tc_c2b, dtype=torch.float32)
    return se3_log_map(T[None].permute(0, 2, 1), eps=eps)[0]
