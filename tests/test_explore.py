"""Host side of the space-exploration scorer (easyhec_b200/explore.py): batched FK -> mvp composition against the
per-candidate composition the reference performs (render_api.py:70-96), and the bookkeeping around the device call."""
import numpy as np
import torch

from easyhec_b200.explore import candidate_mvps, score_candidates, select_next_qpos
from easyhec_b200.scenes import FRANKA_K, SAMPLE_POSE, perturb_pose, scaled_K
from easyhec_b200.urdf_fk import URDFKinematics
from util import mvp_of

URDF = """<robot name="three_link">
  <link name="base"/><link name="l1"/><link name="l2"/><link name="l3"/>
  <joint name="j1" type="revolute"><origin rpy="0 0 0" xyz="0 0 0.3"/><parent link="base"/><child link="l1"/>
    <axis xyz="0 0 1"/><limit lower="-3" upper="3"/></joint>
  <joint name="j2" type="revolute"><origin rpy="-1.57079632679 0 0" xyz="0.1 0 0"/><parent link="l1"/><child link="l2"/>
    <axis xyz="0 0 1"/><limit lower="-2" upper="2"/></joint>
  <joint name="j3" type="prismatic"><origin rpy="0 0.3 0" xyz="0 -0.25 0"/><parent link="l2"/><child link="l3"/>
    <axis xyz="0 1 0"/><limit lower="0" upper="0.1"/></joint>
</robot>"""


def _kin(tmp_path):
    p = tmp_path / "r.urdf"
    p.write_text(URDF)
    return URDFKinematics(str(p))


def test_candidate_mvps_equal_per_candidate_composition(tmp_path):
    kin = _kin(tmp_path)
    rng = np.random.RandomState(0)
    H, W = 1080, 1920
    K = scaled_K(H, W, FRANKA_K)
    q = rng.uniform(-1, 1, size=(5, 2))                      # shorter than the dof: padded on the right
    cams = np.stack([perturb_pose(SAMPLE_POSE, np.random.RandomState(10 + c), 0.05, 5.0) for c in range(3)])
    links = [0, 1, 2, 3]
    mvp = candidate_mvps(kin, links, q, cams, K, H, W, pad_right=1)
    assert mvp.shape == (5, 3, 4, 4, 4) and mvp.dtype == torch.float32 and mvp.is_contiguous()
    qp = np.concatenate([q, np.zeros((5, 1))], 1)
    for qi in range(5):
        lp = kin.forward(qp[qi], links=links).numpy()
        for c in range(3):
            for l in range(4):
                want = mvp_of(K, H, W, cams[c] @ lp[l])     # proj @ flip @ (Tc_c2b @ link_pose), like render_mask
                assert np.allclose(mvp[qi, c, l].numpy(), want, rtol=2e-5, atol=2e-5)


class _FakeCtx:
    device = torch.device("cpu")

    def explore_scores(self, mesh_ids, mvp, H, W):
        self.seen = (list(mesh_ids), tuple(mvp.shape), H, W)
        return torch.arange(mvp.shape[0], dtype=torch.float64) + 1.0


def test_score_candidates_masks_invalid_and_selection_skips_history(tmp_path):
    kin = _kin(tmp_path)
    ctx = _FakeCtx()
    q = np.zeros((4, 3))
    s = score_candidates(ctx, [7, 8, 9, 10], kin, [0, 1, 2, 3], q, SAMPLE_POSE[None], scaled_K(48, 64), 48, 64,
                         valid=[True, True, False, True], device_fk=False)
    assert ctx.seen == ([7, 8, 9, 10], (4, 1, 4, 4, 4), 48, 64)
    assert s.tolist() == [1.0, 2.0, 0.0, 4.0]                 # rejected candidates score 0 (space_explorer.py:109,120,135)
    assert select_next_qpos(s) == (3, 4.0)
    assert select_next_qpos(s, history=[3]) == (1, 2.0)


def test_xarm_chain_urdf_reproduces_the_fixture_kinematics(tmp_path):
    """The URDF that the config-4 GPU test builds from the fixture's joint origins gives the chain's own poses."""
    from easyhec_b200.scenes import chain_fk, load_xarm7
    from util import xarm_urdf
    fx = load_xarm7()
    kin = xarm_urdf(tmp_path, fx)
    assert kin.dof == 7 and kin.link_names == list(fx["names"])
    q = np.random.RandomState(3).uniform(-2, 2, size=(4, 7))
    got = kin.forward(q).numpy()
    for i in range(4):
        assert np.allclose(got[i], chain_fk(fx["joint_origin"], fx["joint_axis"], q[i]), atol=1e-9)


def test_shard_candidates_is_a_block_partition():
    from easyhec_b200.explore import shard_candidates
    for n in (0, 1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            parts = [shard_candidates(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            assert all(parts[r][0] + parts[r][1] == parts[r + 1][0] for r in range(world - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
