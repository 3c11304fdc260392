"""Offline dataset and checkpoint formats of EasyHeC, so that real captures can be replayed through this path.

On-disk dataset (easyhec/data/datasets/xarm_real.py:22-64, docs/franka_offline.md):

    <data_dir>/color/000000.png ...   RGB frames (optional here: the mask loop never reads them)
    <data_dir>/mask/000000.png  ...   annotated robot masks; a pixel is foreground iff ``cv2.imread(path, 2) > 0``
    <data_dir>/qpos/000000.txt  ...   joint positions (np.loadtxt), padded with zeros up to the robot's dof
    <data_dir>/K.txt                  3x3 intrinsics
    <data_dir>/Tc_c2b.txt             optional 4x4 ground-truth camera pose (identity when absent)

``OfflineDataset`` mirrors ``XarmRealDataset``: the same attributes (``masks`` f32 (N,H,W), ``qpos``, ``link_poses``
f32 (N,L,4,4) of ``use_links``, ``K``, ``Tc_c2b``) and the same ``__getitem__`` dictionary.  The link poses come from
``URDFKinematics`` (easyhec_b200/urdf_fk.py) instead of sapien / pinocchio (structures/sapien_kin.py:26-30).

Checkpoints (easyhec/trainer/base.py ``save``: ``{'model': state_dict, 'epoch', 'best_val_loss', 'global_steps'}``
with ``model['dof']`` the 6-vector and ``model['history_ops']`` the (10000, 6) trajectory buffer, rb_solver.py:36-39)
are written in that layout so that the reference's ``tools/validate.py:24-29`` reads them unchanged.
"""
import glob
import os
import os.path as osp

import numpy as np
import torch

from .se3 import dof_to_matrix
from .urdf_fk import URDFKinematics

__all__ = ["OfflineDataset", "write_offline_dataset", "save_checkpoint", "load_checkpoint", "read_mask", "HISTORY_CAPACITY"]

HISTORY_CAPACITY = 10000   # rb_solver.py:39


def read_mask(path: str) -> np.ndarray:
    """bool (H, W): ``cv2.imread(path, 2) > 0`` (xarm_real.py:36); flag 2 = IMREAD_ANYDEPTH, single channel."""
    import cv2
    m = cv2.imread(path, 2)
    if m is None:
        raise FileNotFoundError(path)
    return m > 0


class OfflineDataset(torch.utils.data.Dataset):
    def __init__(self, data_dir: str, urdf_path: str, use_links, ds_len: int = -1, load_color: bool = False):
        self.data_dir = data_dir
        if ds_len < 0:
            ds_len = 1000000
        rgb_paths = sorted(glob.glob(osp.join(data_dir, "color", "*.png")))[:ds_len]
        mask_paths = sorted(glob.glob(osp.join(data_dir, "mask", "*.png")))[:ds_len]
        qpos_paths = sorted(glob.glob(osp.join(data_dir, "qpos", "*.txt")))[:ds_len]
        if not qpos_paths:
            raise FileNotFoundError("no qpos/*.txt under %s" % data_dir)
        self.nimgs = len(rgb_paths) if rgb_paths else len(qpos_paths)
        self.images = []
        if load_color:
            import cv2
            for p in rgb_paths:
                self.images.append(cv2.cvtColor(cv2.imread(p, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB))
        masks = [read_mask(p) for p in mask_paths]
        self.masks = torch.from_numpy(np.stack(masks)).float() if masks else []
        self.kin = URDFKinematics(urdf_path)
        self.use_links = list(use_links)
        self.qpos = [np.atleast_1d(np.loadtxt(p)) for p in qpos_paths]
        q = np.zeros((len(self.qpos), self.kin.dof))
        for i, v in enumerate(self.qpos):   # the reference pads with zeros up to the robot's dof (xarm_real.py:46)
            q[i, :min(len(v), self.kin.dof)] = v[:self.kin.dof]
        self.link_poses = self.kin.forward(q, links=self.use_links).float()
        self.K = torch.from_numpy(np.loadtxt(osp.join(data_dir, "K.txt"))).float()
        tc = osp.join(data_dir, "Tc_c2b.txt")
        self.Tc_c2b = torch.from_numpy(np.loadtxt(tc) if osp.exists(tc) else np.eye(4)).float()

    def __len__(self):
        return self.nimgs

    def __getitem__(self, idx):
        return {"rgb": self.images[idx] if self.images else np.zeros((0,), np.uint8), "qpos": self.qpos[idx],
                "K": self.K, "link_poses": self.link_poses[idx], "Tc_c2b": self.Tc_c2b,
                "mask": self.masks[idx] if len(self.masks) else torch.zeros(0)}

    def batch(self):
        """All views as one batch -- what RBSolver.forward consumes (its assert global_step == 0 demands a single batch)."""
        return {"mask": self.masks, "link_poses": self.link_poses, "K": self.K[None].expand(len(self.qpos), 3, 3),
                "Tc_c2b": self.Tc_c2b[None].expand(len(self.qpos), 4, 4), "global_step": 0}


def write_offline_dataset(data_dir: str, masks, qpos, K, Tc_c2b=None, colors=None):
    """Write the on-disk layout above.  masks (N,H,W) bool / {0,1}; qpos (N,dof); K (3,3); colors (N,H,W,3) u8 RGB."""
    import cv2
    masks = np.asarray(masks)
    for sub in ("mask", "qpos") + (("color",) if colors is not None else ()):
        os.makedirs(osp.join(data_dir, sub), exist_ok=True)
    for i in range(len(masks)):
        cv2.imwrite(osp.join(data_dir, "mask", "%06d.png" % i), (masks[i] > 0).astype(np.uint8) * 255)
        np.savetxt(osp.join(data_dir, "qpos", "%06d.txt" % i), np.asarray(qpos[i], dtype=np.float64))
        if colors is not None:
            cv2.imwrite(osp.join(data_dir, "color", "%06d.png" % i), cv2.cvtColor(np.asarray(colors[i]), cv2.COLOR_RGB2BGR))
    np.savetxt(osp.join(data_dir, "K.txt"), np.asarray(K, dtype=np.float64))
    if Tc_c2b is not None:
        np.savetxt(osp.join(data_dir, "Tc_c2b.txt"), np.asarray(Tc_c2b, dtype=np.float64))


def adam_state_dict(adam_state=None, lr: float = 3e-3, weight_decay: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8):
    """``torch.optim.Adam(...).state_dict()`` of the single 6-vector parameter, from the device solver's state
    ``{m[6], v[6], t}`` (ehb_adam_step) -- what the reference's ``BaseTrainer.load`` hands to
    ``optimizer.load_state_dict`` (trainer/base.py:398).  No state yet: an optimizer that has not stepped."""
    group = {"lr": float(lr), "betas": tuple(betas), "eps": float(eps), "weight_decay": float(weight_decay), "amsgrad": False,
             "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None, "params": [0]}
    state = {}
    if adam_state is not None:
        a = torch.as_tensor(adam_state, dtype=torch.float32).detach().cpu().reshape(13)
        if a[12] > 0:
            state[0] = {"step": torch.tensor(float(a[12])), "exp_avg": a[0:6].clone(), "exp_avg_sq": a[6:12].clone()}
    return {"state": state, "param_groups": [group]}


def adam_state_from_dict(sd):
    """The inverse: an Adam ``state_dict`` -> the device solver's 13 floats."""
    a = torch.zeros(13)
    st = sd.get("state", {}).get(0)
    if st:
        a[0:6] = st["exp_avg"].float().reshape(6); a[6:12] = st["exp_avg_sq"].float().reshape(6); a[12] = float(st["step"])
    return a


def save_checkpoint(path: str, dof, history_ops=None, global_steps: int = 0, epoch: int = 0, best_val_loss: float = 1e10,
                    optimizer_state=None, adam_state=None, lr: float = 3e-3, weight_decay: float = 5e-4, scheduler_state=None):
    """Reference-layout checkpoint (trainer/base.py:380-386): ``ckpt['model']['dof']`` (6,), ``ckpt['model']['history_ops']``
    (10000, 6), ``optimizer`` (Adam state_dict layout, from ``adam_state`` = the device solver's 13 floats unless a ready
    ``optimizer_state`` is given), ``scheduler`` (the constant-lr schedule's counters), ``epoch``, ``best_val_loss``,
    ``global_steps`` -- every key ``BaseTrainer.load`` reads (base.py:388-402)."""
    dof = torch.as_tensor(dof, dtype=torch.float32).detach().cpu().reshape(6)
    hist = torch.zeros(HISTORY_CAPACITY, 6)
    if history_ops is not None:
        h = torch.as_tensor(history_ops, dtype=torch.float32).detach().cpu().reshape(-1, 6)[:HISTORY_CAPACITY]
        hist[:len(h)] = h
    d = {"model": {"dof": dof, "history_ops": hist}, "epoch": int(epoch), "best_val_loss": float(best_val_loss),
         "global_steps": int(global_steps)}
    d["optimizer"] = optimizer_state if optimizer_state is not None else adam_state_dict(adam_state, lr, weight_decay)
    d["scheduler"] = scheduler_state if scheduler_state is not None else {"last_epoch": int(global_steps), "_step_count": int(global_steps) + 1}
    os.makedirs(osp.dirname(osp.abspath(path)), exist_ok=True)
    torch.save(d, path)
    return path


def load_checkpoint(path: str):
    """-> dict(dof (6,), Tc_c2b (4,4) = se3_exp_map(dof)^T as tools/validate.py:27-28 computes it, history_ops (n,6) with
    the unused (all-zero) tail dropped, global_steps, epoch)."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    model = ckpt["model"]
    dof = model["dof"].float().reshape(6)
    hist = model.get("history_ops", torch.zeros(0, 6)).float()
    keep = (hist != 0).any(dim=1)
    n = int(keep.nonzero().max().item()) + 1 if keep.any() else 0
    return {"dof": dof, "Tc_c2b": dof_to_matrix(dof), "history_ops": hist[:n], "global_steps": int(ckpt.get("global_steps", 0)),
            "epoch": int(ckpt.get("epoch", 0)),
            "adam_state": adam_state_from_dict(ckpt["optimizer"]) if isinstance(ckpt.get("optimizer"), dict) else torch.zeros(13)}
