"""Offline dataset / checkpoint formats (easyhec_b200/dataset.py) against the reference's reader semantics
(easyhec/data/datasets/xarm_real.py:22-64) and checkpoint layout (trainer/base.py save, tools/validate.py:24-29)."""
import os

import numpy as np
import pytest
import torch

from easyhec_b200.dataset import (HISTORY_CAPACITY, OfflineDataset, load_checkpoint, read_mask, save_checkpoint,
                                  write_offline_dataset)
from easyhec_b200.se3 import dof_to_matrix, matrix_to_dof

URDF = """<robot name="two_link">
  <link name="base"/><link name="l1"/><link name="l2"/><link name="tool"/>
  <joint name="j1" type="revolute"><origin rpy="0 0 0" xyz="0 0 0.3"/><parent link="base"/><child link="l1"/>
    <axis xyz="0 0 1"/><limit lower="-3" upper="3"/></joint>
  <joint name="j2" type="revolute"><origin rpy="-1.57079632679 0 0" xyz="0.1 0 0"/><parent link="l1"/><child link="l2"/>
    <axis xyz="0 0 1"/><limit lower="-2" upper="2"/></joint>
  <joint name="jt" type="fixed"><origin rpy="0 0 0.5" xyz="0 0 0.2"/><parent link="l2"/><child link="tool"/></joint>
</robot>"""


@pytest.fixture()
def tiny_dataset(tmp_path):
    rng = np.random.RandomState(0)
    masks = rng.rand(3, 24, 32) > 0.6
    qpos = rng.uniform(-1, 1, size=(3, 1))          # shorter than the robot's dof: padded with zeros like the reference
    K = np.array([[30.0, 0, 16], [0, 31.0, 12], [0, 0, 1]])
    Tc = np.eye(4); Tc[:3, 3] = [0.1, -0.2, 1.0]
    d = tmp_path / "data"
    write_offline_dataset(str(d), masks, qpos, K, Tc_c2b=Tc, colors=(rng.rand(3, 24, 32, 3) * 255).astype(np.uint8))
    urdf = tmp_path / "robot.urdf"
    urdf.write_text(URDF)
    return str(d), str(urdf), masks, qpos, K, Tc


def test_roundtrip_matches_reference_reader_semantics(tiny_dataset):
    d, urdf, masks, qpos, K, Tc = tiny_dataset
    ds = OfflineDataset(d, urdf, use_links=[0, 1, 2, 3], load_color=True)
    assert len(ds) == 3
    assert ds.masks.dtype == torch.float32 and np.array_equal(ds.masks.numpy() > 0, masks)       # cv2.imread(path, 2) > 0 -> float
    assert set(np.unique(ds.masks.numpy())) <= {0.0, 1.0}
    assert np.allclose(ds.K.numpy(), K) and np.allclose(ds.Tc_c2b.numpy(), Tc)
    assert ds.link_poses.shape == (3, 4, 4, 4) and ds.link_poses.dtype == torch.float32
    item = ds[1]
    assert set(item) == {"rgb", "qpos", "K", "link_poses", "Tc_c2b", "mask"}                      # xarm_real.py:76-83
    assert item["rgb"].shape == (24, 32, 3)
    # forward kinematics of the padded qpos: j1 rotates about z at height 0.3, j2 = 0
    q = float(qpos[1, 0])
    want = np.eye(4); want[:3, :3] = [[np.cos(q), -np.sin(q), 0], [np.sin(q), np.cos(q), 0], [0, 0, 1]]; want[2, 3] = 0.3
    assert np.allclose(item["link_poses"][1].numpy(), want, atol=1e-6)
    b = ds.batch()
    assert b["mask"].shape == (3, 24, 32) and b["K"].shape == (3, 3, 3) and b["global_step"] == 0


def test_missing_pose_file_defaults_to_identity(tiny_dataset):
    d, urdf, *_ = tiny_dataset
    os.remove(os.path.join(d, "Tc_c2b.txt"))
    ds = OfflineDataset(d, urdf, use_links=[1])
    assert np.array_equal(ds.Tc_c2b.numpy(), np.eye(4, dtype=np.float32))                         # xarm_real.py:58-62


def test_mask_threshold_is_any_nonzero(tmp_path):
    import cv2
    m = np.zeros((4, 5), np.uint8); m[1, 2] = 1; m[3, 4] = 255
    p = str(tmp_path / "m.png")
    cv2.imwrite(p, m)
    assert np.array_equal(read_mask(p), m > 0)


def test_checkpoint_layout_and_roundtrip(tmp_path):
    T = np.eye(4); T[:3, :3] = [[0, -1, 0], [1, 0, 0], [0, 0, 1]]; T[:3, 3] = [0.3, -0.1, 0.9]
    dof = matrix_to_dof(T)
    hist = torch.randn(17, 6)
    p = save_checkpoint(str(tmp_path / "models" / "model_iteration_000017.pth"), dof, hist, global_steps=17, epoch=17)
    raw = torch.load(p, map_location="cpu", weights_only=False)
    assert set(raw) >= {"model", "epoch", "best_val_loss", "global_steps"}                         # trainer/base.py save()
    assert raw["model"]["dof"].shape == (6,) and raw["model"]["history_ops"].shape == (HISTORY_CAPACITY, 6)   # rb_solver.py:36-39
    # what tools/validate.py:27-28 does with it
    Tc = dof_to_matrix(raw["model"]["dof"]).numpy()
    assert np.allclose(Tc, T, atol=1e-5)
    back = load_checkpoint(p)
    assert torch.equal(back["history_ops"], hist) and back["global_steps"] == 17
    assert np.allclose(back["Tc_c2b"].numpy(), T, atol=1e-5)


def test_checkpoint_has_every_key_the_reference_trainer_loads(tmp_path):
    """BaseTrainer.load (trainer/base.py:388-402) reads model, optimizer, scheduler, epoch, best_val_loss, global_steps;
    the optimizer entry must load into a torch.optim.Adam over the single 6-vector parameter."""
    from easyhec_b200.dataset import adam_state_dict
    dof = torch.arange(6, dtype=torch.float32) * 0.1
    st = torch.cat([torch.full((6,), 0.25), torch.full((6,), 0.5), torch.tensor([7.0])])
    p = save_checkpoint(str(tmp_path / "m.pth"), dof, None, global_steps=7, epoch=7, adam_state=st, lr=3e-3, weight_decay=5e-4)
    raw = torch.load(p, map_location="cpu", weights_only=False)
    assert {"model", "optimizer", "scheduler", "epoch", "best_val_loss", "global_steps"} <= set(raw)
    param = torch.nn.Parameter(dof.clone())
    opt = torch.optim.Adam([param], lr=1.0)
    opt.load_state_dict(raw["optimizer"])
    assert opt.param_groups[0]["lr"] == 3e-3 and opt.param_groups[0]["weight_decay"] == 5e-4
    s0 = opt.state[param]
    assert float(s0["step"]) == 7 and torch.allclose(s0["exp_avg"], st[0:6]) and torch.allclose(s0["exp_avg_sq"], st[6:12])
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=10**9)          # (a generic scheduler: __dict__.update of the entry)
    sched.load_state_dict(raw["scheduler"])
    assert torch.allclose(load_checkpoint(p)["adam_state"], st)
    assert adam_state_dict(None)["state"] == {}          # an optimizer that has not stepped
