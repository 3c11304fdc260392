"""Synthetic calibration scenes for tests and benchmarks (host side, numpy).

Real xArm7 link meshes + joint chain come from tests/golden/xarm7_links.npz (made from the reference's
assets by tools/make_fixture_meshes.py); everything else -- joint angles, views, reference masks --
is generated from seeds, as BASELINE.json's configs prescribe.
"""
import os

import numpy as np

from .meshio import Mesh, synthetic_links

__all__ = ["SAMPLE_POSE", "XARM_K", "FRANKA_K", "load_xarm7", "chain_fk", "scaled_K", "perturb_pose", "make_scene",
           "franka_like_links", "FRANKA_FACE_COUNTS", "fit_camera", "onscreen_fraction", "franka_cfg3_scene"]

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


def _scene_points(sc, stride=4):
    """Vertices of every (view, link) in the robot base frame, subsampled: (N, 3) float64."""
    pts = []
    for b in range(sc["link_poses"].shape[0]):
        for l, m in enumerate(sc["meshes"]):
            T = sc["link_poses"][b, l].astype(np.float64)
            v = np.asarray(m.vertices[::stride], np.float64)
            pts.append(v @ T[:3, :3].T + T[:3, 3])
    return np.concatenate(pts)


def onscreen_fraction(sc, H, W, Tc_c2b=None, stride=4):
    """Fraction of the scene's vertices (all views, all links) that project inside the image and lie in front of the camera."""
    T = np.asarray(sc["Tc_c2b"] if Tc_c2b is None else Tc_c2b, np.float64)
    p = _scene_points(sc, stride) @ T[:3, :3].T + T[:3, 3]
    K = np.asarray(sc["K"], np.float64)
    z = np.maximum(p[:, 2], 1e-9)
    u, v = K[0, 0] * p[:, 0] / z + K[0, 2], K[1, 1] * p[:, 1] / z + K[1, 2]
    return float(((p[:, 2] > 0) & (u >= 0) & (u < W) & (v >= 0) & (v < H)).mean())


def fit_camera(sc, H, W, fill=1.0, keep=0.95, iters=60):
    """A camera pose with SAMPLE_POSE's orientation, translated so that the robot of EVERY view of the scene lies inside
    the image (the central `keep` percentile box of the projected vertices fills `fill` of the tighter image dimension,
    so about `keep` of the vertices are on screen).
    One pose for all views, as RBSolver solves for a single Tc_c2b (rb_solver.py:52)."""
    T = np.asarray(sc["Tc_c2b"], np.float64).copy()
    K = np.asarray(sc["K"], np.float64)
    X = _scene_points(sc, 4) @ T[:3, :3].T
    for _ in range(iters):
        p = X + T[:3, 3]
        z = np.maximum(p[:, 2], 1e-3)
        u, v = K[0, 0] * p[:, 0] / z + K[0, 2], K[1, 1] * p[:, 1] / z + K[1, 2]
        q = [50.0 * (1.0 - keep), 100.0 - 50.0 * (1.0 - keep)]
        u0, u1 = np.percentile(u, q); v0, v1 = np.percentile(v, q)
        zm = float(np.median(z))
        T[0, 3] += (W / 2 - (u0 + u1) / 2) * zm / K[0, 0]
        T[1, 3] += (H / 2 - (v0 + v1) / 2) * zm / K[1, 1]
        r = max((u1 - u0) / (fill * W), (v1 - v0) / (fill * H))
        T[2, 3] += zm * (r - 1.0) * 0.7
    return T


def franka_cfg3_scene(H=720, W=1280, views=20):
    """BASELINE.json config 3 on the REAL Franka visual meshes (link0-7 + hand, 133,676 triangles: welded DAE assets of
    configs/franka/example_franka_offline.yaml:9-19, carried in tests/golden/franka_offline.npz), 20 views with
    qpos ~ U(joint limits) (tests/golden/franka_cfg3.npz), K = the xArm default scaled to (H, W), camera = the YAML's
    init_Tc_c2b (configs/franka/example_franka_offline.yaml:5-8)."""
    g = os.path.join(_ROOT, "tests", "golden")
    d = np.load(os.path.join(g, "franka_offline.npz"))
    c = np.load(os.path.join(g, "franka_cfg3.npz"))
    meshes = [Mesh(d[str(n) + "_v"], d[str(n) + "_f"]) for n in d["names"]]
    T = np.asarray(d["yaml_init_Tc_c2b"], np.float64).copy()
    u, _, vt = np.linalg.svd(T[:3, :3])
    T[:3, :3] = u @ vt
    idx = np.arange(views) % len(c["link_poses"])     # more than 20 views: the joint configurations repeat
    return dict(meshes=meshes, link_poses=c["link_poses"][idx].astype(np.float32), K=scaled_K(H, W), Tc_c2b=T,
                qpos=c["qpos"][idx])
