"""Build tests/golden/franka_cfg3.npz: the link poses of BASELINE.json's config 3 (Franka link0-7 + hand, 20 views,
qpos ~ U(joint limits of assets/franka/urdf/franka.urdf), fingers 0, seed 0 -- SURVEY.md 8d) from the reference's URDF.
Run in the build container only (/root/reference is not present on the GPU box); the meshes themselves travel in
tests/golden/franka_offline.npz (tools/make_fixture_franka.py)."""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from easyhec_b200.urdf_fk import URDFKinematics  # noqa: E402

REF = os.environ.get("EHB_REFERENCE", "/root/reference")
USE_LINKS = [0, 1, 2, 3, 4, 5, 6, 7, 9]       # configs/franka/example_franka_offline.yaml:39


def main():
    kin = URDFKinematics(os.path.join(REF, "assets/franka/urdf/franka.urdf"))
    lim = kin.joint_limits
    rng = np.random.RandomState(0)
    q = rng.uniform(lim[:, 0], lim[:, 1], size=(20, len(lim)))
    q[:, 7:] = 0.0                           # fingers closed, like the 9-dim qpos files of the offline example
    lp = kin.forward(q, links=USE_LINKS).numpy().astype(np.float32)
    out = os.path.join(ROOT, "tests", "golden", "franka_cfg3.npz")
    np.savez_compressed(out, qpos=q, link_poses=lp, joint_limits=lim)
    print("wrote", out, lp.shape, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
