"""Build tests/golden/xarm7_links.npz from the reference's xArm7 assets (run in the build container only).

Inputs (read-only, /root/reference is NOT present on the GPU box, hence this fixture):
  assets/xarm_description/meshes/xarm7/visual/link_base.STL, link1..7.STL   (configs/xarm7/example.yaml:16-22)
  assets/xarm7_with_gripper_reduced_dof.urdf                                  (joint origins + limits)
Output: welded link meshes (vertices f32, faces i32), the URDF joint chain of link_base..link7 as
(origin 4x4, axis) per joint, and the joint limits -- everything bench.py / tests need to pose the arm.
Vertices are stored as float16-exact? No: stored exactly (f32) and compressed; ~0.5 MB.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from easyhec_b200.meshio import load_mesh  # noqa: E402
from easyhec_b200.urdf_fk import URDFKinematics  # noqa: E402

REF = os.environ.get("EHB_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "xarm7_links.npz")


def main():
    names = ["link_base"] + ["link%d" % i for i in range(1, 8)]
    arrs = {"names": np.array(names)}
    for n in names:
        m = load_mesh(os.path.join(REF, "assets/xarm_description/meshes/xarm7/visual", n + ".STL"))
        arrs[n + "_v"] = m.vertices
        arrs[n + "_f"] = m.faces
        print(n, m)
    kin = URDFKinematics(os.path.join(REF, "assets/xarm7_with_gripper_reduced_dof.urdf"))
    chain = []
    for n in names[1:]:
        j = kin._parent_joint[n]
        assert j["parent"] == (names[names.index(n) - 1]), (j["parent"], n)
        chain.append(j)
    arrs["joint_origin"] = np.stack([j["origin"] for j in chain]).astype(np.float64)
    arrs["joint_axis"] = np.stack([j["axis"] for j in chain]).astype(np.float64)
    arrs["joint_limits"] = np.array([[j["lower"], j["upper"]] for j in chain], dtype=np.float64)
    # zero-pose FK of the 8 links, as a cross-check for the FK restatement
    arrs["fk_zero"] = kin.forward(np.zeros(kin.dof), links=names).numpy()
    np.savez_compressed(OUT, **arrs)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
