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


def xarm_urdf(tmp_path, fx):
    """URDF of the xArm7 serial chain rebuilt from the fixture's joint origins / axes (link_base, link1..link7) ->
    URDFKinematics.  rpy is fixed-axis XYZ: R = Rz(y) Ry(p) Rx(r)."""
    from easyhec_b200.urdf_fk import URDFKinematics
    names = fx["names"]
    out = ['<robot name="xarm7_chain">'] + ['<link name="%s"/>' % n for n in names]
    for i in range(len(names) - 1):
        T = np.asarray(fx["joint_origin"][i], np.float64)
        R = T[:3, :3]
        p = np.arcsin(-R[2, 0])
        r = np.arctan2(R[2, 1], R[2, 2])
        y = np.arctan2(R[1, 0], R[0, 0])
        lo, hi = fx["joint_limits"][i]
        out.append('<joint name="joint%d" type="revolute"><origin rpy="%.17g %.17g %.17g" xyz="%.17g %.17g %.17g"/>'
                   '<parent link="%s"/><child link="%s"/><axis xyz="%.17g %.17g %.17g"/><limit lower="%.17g" upper="%.17g"/>'
                   '</joint>' % ((i + 1, r, p, y) + tuple(T[:3, 3]) + (names[i], names[i + 1]) + tuple(fx["joint_axis"][i]) + (lo, hi)))
    out.append("</robot>")
    path = tmp_path / "xarm7_chain.urdf"
    path.write_text("\n".join(out))
    return URDFKinematics(str(path))
