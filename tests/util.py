"""Shared helpers of the parity tests: scenes as plain numpy arrays, fed identically to oracle and CUDA."""
import numpy as np
import torch

from easyhec_b200.meshio import Mesh, concat_meshes
from easyhec_b200.projection import K_to_projection, opencv2gl
from easyhec_b200.scenes import SAMPLE_POSE, make_scene, scaled_K


def mvp_of(K, H, W, pose):
    """fp32 mvp = proj @ flip @ pose, composed on the CPU exactly like the product composes it."""
    K = torch.as_tensor(np.asarray(K), dtype=torch.float32)
    pose = torch.as_tensor(np.asarray(pose), dtype=torch.float32)
    return (K_to_projection(K, H, W) @ (opencv2gl() @ pose)).numpy().astype(np.float32)


def scene_mvps(sc, H, W, Tc_c2b=None):
    """(B,L,4,4) fp32 mvps of a make_scene() scene under camera pose Tc_c2b."""
    T = torch.as_tensor(sc["Tc_c2b"] if Tc_c2b is None else Tc_c2b, dtype=torch.float32)
    lp = torch.as_tensor(sc["link_poses"], dtype=torch.float32)
    P = K_to_projection(torch.as_tensor(sc["K"]), H, W) @ opencv2gl()
    return (P @ (T @ lp)).numpy().astype(np.float32)


def zero_pose_robot(xarm):
    """All 8 xArm7 links at qpos = 0 packed into one mesh (the content of assets/xarm7_zeropos.ply)."""
    return concat_meshes([m.transformed(T) for m, T in zip(xarm["meshes"], xarm["fk_zero"])])


def quad(z=1.0, half=5.0):
    """Two huge triangles facing the camera (cover the whole screen)."""
    v = np.array([[-half, -half, z], [half, -half, z], [half, half, z], [-half, half, z]], np.float32)
    f = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    return Mesh(v, f)


def to_dev(a, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
