"""Build tests/golden/franka_offline.npz from the reference's real Franka offline example (run in the build container
only; /root/reference is NOT present on the GPU box, hence this fixture).

Inputs (read-only):
  assets/franka_offline_example.zip            10 views 640x480: mask/*.png, qpos/*.txt, K.txt   (docs/franka_offline.md)
  assets/franka/franka_description/meshes/visual/{link0..7,hand}.dae     (configs/franka/example_franka_offline.yaml:9-19)
  assets/franka/urdf/franka.urdf, use_links [0..7, 9]                      (configs/franka/example_franka_offline.yaml:39-40)
  init_Tc_c2b                                                              (configs/franka/example_franka_offline.yaml:5-8)
Output: welded link meshes, the 10 reference masks (bit-packed), qpos, K, the link poses from the URDF forward
kinematics, the initial pose, and the trajectory of an ORACLE-DRIVEN solve (the CPU restatement in oracle/ under
torch.optim.Adam(lr 3e-3, weight_decay 5e-4), the reference's optimiser: solver/build.py:12-29) that the GPU solver
is compared with in tests/test_gpu_solver.py.
"""
import glob
import io
import os
import sys
import time
import zipfile

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from easyhec_b200.meshio import load_mesh  # noqa: E402
from easyhec_b200.rb_solver import compose_link_mvp  # noqa: E402
from easyhec_b200.se3 import dof_to_matrix, matrix_to_dof, se3_exp_map  # noqa: E402
from easyhec_b200.urdf_fk import URDFKinematics  # noqa: E402
from oracle import oracle  # noqa: E402

REF = os.environ.get("EHB_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden", "franka_offline.npz")
NAMES = ["link0", "link1", "link2", "link3", "link4", "link5", "link6", "link7", "hand"]
USE_LINKS = [0, 1, 2, 3, 4, 5, 6, 7, 9]
INIT = np.array([[9.3969262e-01, 3.4202009e-01, 6.4914198e-09, -6.4085639e-01],
                 [1.7101002e-01, -4.6984622e-01, -8.6602539e-01, 4.9582830e-01],
                 [-2.9619810e-01, 8.1379771e-01, -4.9999991e-01, 1.2412001e+00],
                 [0.0, 0.0, 0.0, 1.0]])
# The YAML pose overlaps the annotated masks with IoU 0.25 and the oracle-driven solve drifts away from it; a random
# search over poses (mean per-view IoU as the score, union_binary of the oracle as the renderer, this session) ends at
# the pose below (mean IoU 0.70; single views fitted alone reach 0.79-0.93).  tools/franka_sync_probe.py shows why: the
# capture's images lag its joint positions by about one sample in four of the ten views (IoU 0.87 once that is modelled).  The fixture's solve starts 2 cm / 2 deg away from it, like a
# hand-tuned initialisation (tools/manual_tune_franka_init.py in the reference).
TUNED = np.array([[0.96903512, -0.24688871, -0.00410332, -0.47387618],
                  [-0.12451614, -0.47423916, -0.87154636, 0.52374082],
                  [0.21322902, 0.84507002, -0.49029604, 0.96514138],
                  [0.0, 0.0, 0.0, 1.0]])
H, W = 480, 640


def oracle_solve(meshes, link_poses, K, masks, init, iters, record):
    packed = oracle.pack_links(meshes)
    dof = torch.nn.Parameter(matrix_to_dof(torch.as_tensor(init, dtype=torch.float32)).clone())
    opt = torch.optim.Adam([dof], lr=3e-3, weight_decay=5e-4)
    lp = torch.as_tensor(link_poses, dtype=torch.float32)
    Kt = torch.as_tensor(K, dtype=torch.float32)
    ref = masks.astype(np.float32)
    traj, losses = {}, {}
    for it in range(iters + 1):
        if it in record:
            traj[it] = dof.detach().numpy().copy()
        if it == iters:
            break
        opt.zero_grad()
        Tc = se3_exp_map(dof[None]).permute(0, 2, 1)[0]
        mvp = compose_link_mvp(Kt, H, W, Tc, lp)
        out = oracle.render_views(packed, mvp.detach().numpy().astype(np.float32), ref, H, W)
        loss = float(out["loss_per_view"].sum() / len(ref))
        mvp.backward(torch.as_tensor(out["g_mvp"], dtype=torch.float32))
        opt.step()
        if it in record or it % 50 == 0:
            losses[it] = loss
            print("iter %4d loss %.2f" % (it, loss), flush=True)
    return traj, losses


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    from easyhec_b200.scenes import perturb_pose
    u, _, vt = np.linalg.svd(TUNED[:3, :3])
    tuned = TUNED.copy(); tuned[:3, :3] = u @ vt
    start = perturb_pose(tuned, np.random.RandomState(3), 0.02, 2.0)
    meshes = [load_mesh(os.path.join(REF, "assets/franka/franka_description/meshes/visual", n + ".dae")) for n in NAMES]
    z = zipfile.ZipFile(os.path.join(REF, "assets/franka_offline_example.zip"))
    import cv2
    mask_names = sorted(n for n in z.namelist() if "/mask/" in n and n.endswith(".png"))
    qpos_names = sorted(n for n in z.namelist() if "/qpos/" in n and n.endswith(".txt"))
    masks = np.stack([cv2.imdecode(np.frombuffer(z.read(n), np.uint8), 2) > 0 for n in mask_names])   # cv2.imread(path, 2) > 0
    qpos = np.stack([np.loadtxt(io.BytesIO(z.read(n))) for n in qpos_names])
    K = np.loadtxt(io.BytesIO(z.read([n for n in z.namelist() if n.endswith("K.txt")][0])))
    kin = URDFKinematics(os.path.join(REF, "assets/franka/urdf/franka.urdf"))
    link_poses = kin.forward(qpos, links=USE_LINKS).numpy()
    assert masks.shape == (10, H, W) and link_poses.shape == (10, 9, 4, 4)
    record = sorted(set([0, 1, 2, 5, 10, 20, 50, 100, 200, iters]))
    t0 = time.time()
    traj, losses = oracle_solve(meshes, link_poses, K, masks, start, iters, record)
    print("oracle-driven solve: %d iterations in %.0f s" % (iters, time.time() - t0))
    arrs = {"names": np.array(NAMES), "masks_packed": np.packbits(masks, axis=-1), "qpos": qpos, "K": K,
            "link_poses": link_poses.astype(np.float32), "yaml_init_Tc_c2b": INIT, "tuned_Tc_c2b": tuned, "start_Tc_c2b": start,
            "H": H, "W": W,
            "traj_iters": np.array(sorted(traj)), "traj_dof": np.stack([traj[k] for k in sorted(traj)]),
            "loss_iters": np.array(sorted(losses)), "loss_values": np.array([losses[k] for k in sorted(losses)])}
    for n, m in zip(NAMES, meshes):
        arrs[n + "_v"] = m.vertices
        arrs[n + "_f"] = m.faces
    np.savez_compressed(OUT, **arrs)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
    Tc = dof_to_matrix(torch.as_tensor(traj[iters]))
    print("final Tc_c2b\n", Tc.numpy())


if __name__ == "__main__":
    main()
