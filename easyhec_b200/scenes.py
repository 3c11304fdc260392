"""Synthetic calibration scenes for tests and benchmarks (host side, numpy).

Real xArm7 link meshes + joint chain come from tests/golden/xarm7_links.npz (made from the reference's
assets by tools/make_fixture_meshes.py); everything else -- joint angles, views, reference masks --
is generated from seeds, as BASELINE.json's configs prescribe.
"""
import os

import numpy as np

from .meshio import Mesh, synthetic_links

__all__ = ["SAMPLE_POSE", "XARM_K", "FRANKA_K", "load_xarm7", "chain_fk", "scaled_K", "perturb_pose", "make_scene",
           "franka_like_links", "FRANKA_FACE_COUNTS"]

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# camera-from-base pose used by the reference's own renderer demo (nvdiffrast_renderer.py:77-80)
SAMPLE_POSE = np.array([[0.99638397, -0.0846324, 0.00750877, -0.20668708],
                        [-0.00875172, -0.19013488, -0.9817189, 0.08405855],
                        [0.0845129, 0.97810328, -0.19018805, 0.77892876],
                        [0., 0., 0., 1.]], dtype=np.float64)
# default intrinsics at 1280x720 (config/defaults.py:14-16) and 1920x1080 (config/defaults_franka.py:14-16)
XARM_K = (906.805, 906.680, 650.198, 367.714, 1280, 720)
FRANKA_K = (1352.21, 1352.43, 963.35, 529.40, 1920, 1080)
# per-link triangle counts of the Franka visual meshes
FRANKA_FACE_COUNTS = (20483, 12516, 12716, 14233, 14621, 18327, 21620, 12082, 7078)


def scaled_K(H, W, base=XARM_K):
    fx, fy, cx, cy, W0, H0 = base
    sx, sy = W / W0, H / H0
    return np.array([[fx * sx, 0, cx * sx], [0, fy * sy, cy * sy], [0, 0, 1]], dtype=np.float32)


def load_xarm7(path=None):
    """-> dict(names, meshes [link_base, link1..link7], joint_origin (7,4,4), joint_axis (7,3), joint_limits (7,2))"""
    path = path or os.path.join(_ROOT, "tests", "golden", "xarm7_links.npz")
    d = np.load(path)
    names = [str(n) for n in d["names"]]
    return dict(names=names, meshes=[Mesh(d[n + "_v"], d[n + "_f"]) for n in names],
                joint_origin=d["joint_origin"], joint_axis=d["joint_axis"], joint_limits=d["joint_limits"],
                fk_zero=d["fk_zero"])


def _axis_rot(axis, q):
    x, y, z = axis
    Kx = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
    return np.eye(3) + np.sin(q) * Kx + (1 - np.cos(q)) * (Kx @ Kx)


def chain_fk(joint_origin, joint_axis, q):
    """Serial revolute chain: poses (n+1,4,4) of the base and the n child links in the base frame."""
    T = np.eye(4)
    out = [T.copy()]
    for i in range(len(joint_origin)):
        M = np.eye(4)
        M[:3, :3] = _axis_rot(joint_axis[i], q[i] if i < len(q) else 0.0)
        T = T @ joint_origin[i] @ M
        out.append(T.copy())
    return np.stack(out)


def perturb_pose(T, rng, trans=0.03, rot_deg=3.0):
    """T @ exp(delta): delta_t ~ U(+-trans), delta_w ~ U(+-rot_deg)."""
    dt = rng.uniform(-trans, trans, 3)
    w = np.deg2rad(rng.uniform(-rot_deg, rot_deg, 3))
    th = np.linalg.norm(w) + 1e-30
    Kx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]) / th
    D = np.eye(4)
    D[:3, :3] = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * (Kx @ Kx)
    D[:3, 3] = dt
    return T @ D


def franka_like_links(seed=0):
    """Procedural stand-ins with the Franka's per-link triangle counts (the DAE assets are not shipped)."""
    return synthetic_links(FRANKA_FACE_COUNTS, radius=0.055, length=0.18, seed=seed)


def make_scene(B, H, W, links="xarm7", seed=0, K_base=XARM_K, joint_range=0.6):
    """-> dict(meshes [L], link_poses (B,L,4,4) f32, K (3,3) f32, Tc_c2b (4,4) f64 ground truth, qpos (B,n)).

    links: "xarm7" = link1..link7 (configs/xarm7/example.yaml:16-22), "xarm7_all" = link_base + link1..7,
    "franka_like" = 9 procedural links on the xArm chain geometry (extra links ride on the last joints)."""
    rng = np.random.RandomState(seed)
    fx = load_xarm7()
    lim = fx["joint_limits"]
    lo = np.maximum(lim[:, 0], -np.pi) * joint_range
    hi = np.minimum(lim[:, 1], np.pi) * joint_range
    qpos = rng.uniform(lo, hi, size=(B, len(lim)))
    poses = np.stack([chain_fk(fx["joint_origin"], fx["joint_axis"], q) for q in qpos])   # (B,8,4,4)
    if links == "xarm7":
        meshes, idx = fx["meshes"][1:], list(range(1, 8))
    elif links == "xarm7_all":
        meshes, idx = fx["meshes"], list(range(0, 8))
    elif links == "franka_like":
        meshes, idx = franka_like_links(seed), [0, 1, 2, 3, 4, 5, 6, 7, 7]
    else:
        raise ValueError(links)
    link_poses = poses[:, idx].copy()
    if links == "franka_like":   # put the 9th link a little further along the last link's axis
        off = np.eye(4)
        off[2, 3] = 0.12
        link_poses[:, 8] = link_poses[:, 8] @ off
    return dict(meshes=meshes, link_poses=link_poses.astype(np.float32), K=scaled_K(H, W, K_base),
                Tc_c2b=SAMPLE_POSE.copy(), qpos=qpos)
