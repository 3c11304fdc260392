"""URDF forward kinematics in PyTorch (batched, differentiable) -- host side.

Stands in for sapien/pinocchio, which the reference uses to turn ``qpos`` into per-link poses
(easyhec/structures/sapien_kin.py:26-30, easyhec/data/datasets/xarm_real.py:42-56,
easyhec/utils/render_api.py:145-192).  Conventions reproduced:

* links are numbered in URDF document order (what `robot.get_links()` yields for these URDFs),
  so xArm ``use_links=[2..8]`` is link1..link7 and ``i+1`` in render_api.py:153 is link_base..link7;
* ``qpos`` lists the movable (revolute / continuous / prismatic) joints in document order; a short
  qpos is zero-padded (xarm_real.py:46-47);
* joint origin rpy is fixed-axis XYZ, i.e. ``R = Rz(y) Ry(p) Rx(r)``.
"""
import xml.etree.ElementTree as ET

import numpy as np
import torch

__all__ = ["URDFKinematics"]


def _rpy_matrix(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


class URDFKinematics:
    def __init__(self, urdf_path: str):
        root = ET.parse(urdf_path).getroot()
        self.link_names = [l.get("name") for l in root.findall("link")]
        self.joints = []
        for j in root.findall("joint"):
            o = j.find("origin")
            xyz = np.array((o.get("xyz", "0 0 0") if o is not None else "0 0 0").split(), dtype=np.float64)
            rpy = np.array((o.get("rpy", "0 0 0") if o is not None else "0 0 0").split(), dtype=np.float64)
            T = np.eye(4)
            T[:3, :3] = _rpy_matrix(*rpy)
            T[:3, 3] = xyz
            ax = j.find("axis")
            axis = np.array((ax.get("xyz") if ax is not None else "1 0 0").split(), dtype=np.float64)
            lim = j.find("limit")
            mimic = j.find("mimic")
            self.joints.append(dict(
                name=j.get("name"), type=j.get("type"), parent=j.find("parent").get("link"),
                child=j.find("child").get("link"), origin=T, axis=axis / (np.linalg.norm(axis) + 1e-30),
                lower=float(lim.get("lower", "0")) if lim is not None else 0.0,
                upper=float(lim.get("upper", "0")) if lim is not None else 0.0,
                mimic=None if mimic is None else (mimic.get("joint"), float(mimic.get("multiplier", "1")),
                                                  float(mimic.get("offset", "0")))))
        self.movable = [j for j in self.joints if j["type"] in ("revolute", "continuous", "prismatic")
                        and j["mimic"] is None]
        self.dof = len(self.movable)
        self._parent_joint = {j["child"]: j for j in self.joints}
        children = set(self._parent_joint)
        self.root_link = [n for n in self.link_names if n not in children][0]

    @property
    def joint_limits(self) -> np.ndarray:
        return np.array([[j["lower"], j["upper"]] for j in self.movable])

    def link_index(self, name: str) -> int:
        return self.link_names.index(name)

    def _joint_motion(self, j, q):
        """q (N,) -> (N,4,4) motion of a movable joint about/along its axis."""
        N = q.shape[0]
        ax = torch.as_tensor(j["axis"], dtype=q.dtype, device=q.device)
        M = torch.eye(4, dtype=q.dtype, device=q.device).repeat(N, 1, 1)
        if j["type"] == "prismatic":
            M[:, :3, 3] = q[:, None] * ax[None]
            return M
        x, y, z = ax
        Kx = torch.stack([torch.zeros(()).to(q), -z, y, z, torch.zeros(()).to(q), -x, -y, x,
                          torch.zeros(()).to(q)]).reshape(3, 3)
        s, c = torch.sin(q)[:, None, None], torch.cos(q)[:, None, None]
        M[:, :3, :3] = torch.eye(3, dtype=q.dtype, device=q.device)[None] + s * Kx[None] + (1 - c) * (Kx @ Kx)[None]
        return M

    def forward(self, qpos, links=None) -> torch.Tensor:
        """qpos (N,<=dof) or (<=dof,) -> poses (N, len(links), 4, 4) of link frames in the root frame."""
        q = torch.as_tensor(qpos)
        if not q.is_floating_point():
            q = q.double()
        single = q.ndim == 1
        if single:
            q = q[None]
        if q.shape[1] < self.dof:
            q = torch.cat([q, q.new_zeros(q.shape[0], self.dof - q.shape[1])], 1)
        qmap = {j["name"]: q[:, i] for i, j in enumerate(self.movable)}
        for j in self.joints:
            if j["mimic"] is not None and j["mimic"][0] in qmap:
                qmap[j["name"]] = qmap[j["mimic"][0]] * j["mimic"][1] + j["mimic"][2]
        N = q.shape[0]
        cache = {self.root_link: torch.eye(4, dtype=q.dtype, device=q.device).repeat(N, 1, 1)}

        def pose(name):
            if name in cache:
                return cache[name]
            j = self._parent_joint[name]
            T = pose(j["parent"]) @ torch.as_tensor(j["origin"], dtype=q.dtype, device=q.device)[None]
            if j["name"] in qmap:
                T = T @ self._joint_motion(j, qmap[j["name"]])
            cache[name] = T
            return T

        if links is None:
            links = range(len(self.link_names))
        out = torch.stack([pose(self.link_names[i] if not isinstance(i, str) else i) for i in links], 1)
        return out[0] if single else out

    __call__ = forward
