"""Known-answer tests that pin the CPU oracle (oracle/ehb_oracle.c) itself.

nvdiffrast is not vendored in the reference tree, so there are no upstream golden vectors for the
rasterize / antialias arithmetic ("parity unpinned").  These tests pin the restatement by
construction: analytic coverage counts, analytic blend weights, gradients against finite differences of
the oracle's own forward, and the FK / mesh fixtures against the reference's assets.
"""
import numpy as np
import pytest

from oracle import oracle
from easyhec_b200.meshio import Mesh, concat_meshes, weld
from easyhec_b200.scenes import SAMPLE_POSE, chain_fk, load_xarm7, make_scene, scaled_K
from util import mvp_of, quad, scene_mvps, zero_pose_robot

H = W = 64
K64 = np.array([[64.0, 0, 32.0], [0, 64.0, 32.0], [0, 0, 1]], np.float32)   # 1 unit at z = 1 -> 64 px


def square(x0, y0, x1, y1, z=1.0):
    """axis-aligned rectangle given in OpenCV pixel coordinates (u right, v down)"""
    a = lambda u: (u - 32) / 64.0 * z
    v = np.array([[a(x0), a(y0), z], [a(x1), a(y0), z], [a(x1), a(y1), z], [a(x0), a(y1), z]], np.float32)
    return v, np.array([[0, 1, 2], [0, 2, 3]], np.int32)


@pytest.mark.parametrize("off,size", [(0.0, 8), (0.5, 8), (0.25, 4), (0.5, 1), (0.75, 3)])
def test_square_coverage_counts(off, size):
    mvp = mvp_of(K64, H, W, np.eye(4))
    v, f = square(10 + off, 20 + off, 10 + off + size, 20 + off + size)
    m = oracle.render_mask(v, f, mvp, H, W, anti_aliasing=False)
    assert m.sum() == size * size            # a shared diagonal / an on-centre edge is owned exactly once
    ys, xs = np.nonzero(m)
    assert xs.max() - xs.min() == size - 1 and ys.max() - ys.min() == size - 1
    # output rows are image rows (row 0 = top): the square sits at v ~ 20, not at H - 20
    assert 19 <= ys.min() <= 21


def test_both_fill_rules_tile_the_plane():
    mvp = mvp_of(K64, H, W, np.eye(4))
    v, f = square(8.5, 8.5, 24.5, 24.5)
    for rule in (0, 1):
        tid, _, _ = oracle.rasterize(oracle.transform(v, mvp), f, H, W, rule=rule)
        assert (tid >= 0).sum() == 16 * 16
        assert set(np.unique(tid)) == {-1, 0, 1}   # both triangles of the quad own pixels, none twice


def test_antialias_weight_of_a_vertical_edge():
    """Right edge of a square at u = 20.3: pixel 19 (centre 19.5) is covered, pixel 20 (centre 20.5) is not.
    The edge crosses the segment between the two centres at 0.8 of the way: the empty pixel receives 0.8 - 0.5,
    and the mask stays within [0, 1]."""
    mvp = mvp_of(K64, H, W, np.eye(4))
    v, f = square(10.0, 10.0, 20.3, 30.0)
    aa = oracle.render_mask(v, f, mvp, H, W, anti_aliasing=True)
    row = aa[20]
    assert row[18] == 1.0
    assert abs(row[20] - 0.3) < 1e-4 and row[19] == 1.0 and row[21] == 0.0
    v, f = square(10.0, 10.0, 19.8, 30.0)        # crossing at 0.3: the covered pixel loses 0.5 - 0.3
    row = oracle.render_mask(v, f, mvp, H, W, anti_aliasing=True)[20]
    assert abs(row[19] - 0.8) < 1e-4 and row[20] == 0.0
    assert aa.min() >= 0.0 and aa.max() <= 1.0


def test_antialias_gradient_matches_finite_differences():
    """d/dx of sum(dy * mask) for a square translated in x; the antialiased mask is piecewise linear in the edge
    position, so central differences of the oracle's own forward are exact up to fp32 noise."""
    v, f = square(10.2, 12.3, 21.4, 30.6)
    rng = np.random.RandomState(0)
    dy = rng.rand(H, W).astype(np.float32)
    pose = np.eye(4)
    mvp = mvp_of(K64, H, W, pose)
    _, st = oracle.render_mask(v, f, mvp, H, W, anti_aliasing=True, save=True)
    _, g_mvp = oracle.render_mask_bwd(v, f, mvp, H, W, st, dy)
    eps = 2e-4

    def loss(dx, dyy):
        p = pose.copy(); p[0, 3] += dx; p[1, 3] += dyy
        m = oracle.render_mask(v, f, mvp_of(K64, H, W, p), H, W, anti_aliasing=True)
        return float((m.astype(np.float64) * dy).sum())

    for axis, d in ((0, (eps, 0.0)), (1, (0.0, eps))):
        fd = (loss(*d) - loss(-d[0], -d[1])) / (2 * eps)
        # chain rule: mvp = P @ pose, d mvp / d pose[axis,3] = P[:, axis] in column 3
        P = mvp_of(K64, H, W, np.eye(4)).astype(np.float64)
        an = float((g_mvp[:, 3] * P[:, axis]).sum())
        assert abs(fd - an) < 2e-2 * max(1.0, abs(fd)), (axis, fd, an)


def test_union_binary_equals_packed_mesh_render(xarm):
    Hh, Ww = 96, 128
    sc = make_scene(2, Hh, Ww, links="xarm7_all", seed=4)
    mvp = scene_mvps(sc, Hh, Ww)
    packed = oracle.pack_links(sc["meshes"])
    got = oracle.union_binary(packed, mvp, Hh, Ww)
    for b in range(2):
        world = concat_meshes([m.transformed(T) for m, T in zip(sc["meshes"], sc["link_poses"][b])])
        want = oracle.render_mask(world.vertices, world.faces, mvp_of(sc["K"], Hh, Ww, sc["Tc_c2b"]), Hh, Ww,
                                  anti_aliasing=False)
        # same geometry posed on the host in fp64 then rounded vs posed by the fp32 mvp: silhouettes agree
        assert (got[b] != want).mean() < 2e-3
        assert got[b].sum() > 50


def test_render_views_composition_and_loss():
    """S = min(sum_l m_l, 1), loss = mean_b sum (S - ref)^2 (rb_solver.py:68-72), from single-link renders."""
    Hh, Ww, B = 60, 80, 2
    sc = make_scene(B, Hh, Ww, links="xarm7", seed=2)
    mvp = scene_mvps(sc, Hh, Ww)
    packed = oracle.pack_links(sc["meshes"])
    ref = (np.random.RandomState(0).rand(B, Hh, Ww) > 0.7).astype(np.float32)
    out = oracle.render_views(packed, mvp, ref, Hh, Ww)
    for b in range(B):
        s = np.zeros((Hh, Ww), np.float32)
        for l, m in enumerate(sc["meshes"]):
            s = s + oracle.render_mask(m.vertices, m.faces, mvp[b, l], Hh, Ww, anti_aliasing=True)
        S = np.minimum(s, 1.0)
        assert np.array_equal(out["masks"][b], S)
        assert np.isclose(out["loss_per_view"][b], ((S - ref[b]).astype(np.float32) ** 2).astype(np.float64).sum(),
                          rtol=1e-12)
    assert np.isclose(out["loss"], out["loss_per_view"].mean())
    assert np.all(out["g_mvp"][:, :, 2, :] == 0)     # nothing flows through clip-space z


def test_variance_scores_match_numpy():
    rng = np.random.RandomState(3)
    m = rng.rand(3, 5, 12, 17) > 0.5
    want = m.astype(np.float64).var(axis=1, ddof=1).reshape(3, -1).sum(1)
    assert np.allclose(oracle.variance_scores(m), want, rtol=1e-12)


def test_empty_and_offscreen_inputs():
    mvp = mvp_of(K64, H, W, np.eye(4))
    m = oracle.render_mask(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.int32), mvp, H, W, anti_aliasing=True)
    assert m.shape == (H, W) and not m.any()
    q = quad(z=-3.0)     # behind the camera
    assert not oracle.render_mask(q.vertices, q.faces, mvp, H, W, anti_aliasing=False).any()
    q = quad(z=1.0)      # covers everything: no silhouette inside the image, AA mask is exactly 1
    assert (oracle.render_mask(q.vertices, q.faces, mvp, H, W, anti_aliasing=True) == 1.0).all()


def test_fk_fixture_reproduces_zero_pose(xarm):
    fk = chain_fk(xarm["joint_origin"], xarm["joint_axis"], np.zeros(7))
    assert np.allclose(fk, xarm["fk_zero"], atol=1e-12)
    robot = zero_pose_robot(xarm)
    lo, hi = robot.vertices.min(0), robot.vertices.max(0)
    # bbox of assets/xarm7_zeropos.ply
    assert np.allclose(lo, [-0.0923, -0.1069, 0.0], atol=2e-3) and np.allclose(hi, [0.2435, 0.1190, 0.6028], atol=2e-3)
    assert len(robot.faces) == 41096 and len(robot.vertices) == 20525


def test_weld_and_adjacency():
    tri = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[1, 0, 0], [1, 1, 0], [0, 1, 0]]], np.float32)
    m = weld(tri)
    assert len(m.vertices) == 4 and m.faces.tolist() == [[0, 1, 2], [1, 3, 2]]
    opp = oracle.build_adjacency(m.faces, 4)
    # shared edge (1,2): opposite vertices are 3 and 0; boundary edges have none
    assert opp[0].tolist() == [3, -1, -1] and opp[1].tolist() == [-1, 0, -1]


def test_config1_plumbing_zero_pose_mask_and_iou(xarm):
    """BASELINE config 1: xArm7 zero pose (the content of assets/xarm7_zeropos.ply: 20,525 vertices /
    41,096 triangles), one view 128x128, K = (90.68, 90.67, 64, 64), Tc_c2b = the sample pose of
    nvdiffrast_renderer.py:77-80; CPU only: mask, mask-L2 loss and IoU against a displaced camera."""
    m = zero_pose_robot(xarm)
    assert len(m.faces) == 41096
    Hh = Ww = 128
    K = np.array([[90.68, 0, 64.0], [0, 90.67, 64.0], [0, 0, 1]], np.float32)
    ref = oracle.render_mask(m.vertices, m.faces, mvp_of(K, Hh, Ww, SAMPLE_POSE), Hh, Ww, anti_aliasing=False)
    assert 0.03 < ref.mean() < 0.5 and ref[:, 0].sum() == 0 and ref[:, -1].sum() == 0      # the arm, fully in view
    aa = oracle.render_mask(m.vertices, m.faces, mvp_of(K, Hh, Ww, SAMPLE_POSE), Hh, Ww, anti_aliasing=True)
    assert aa.min() >= 0.0 and aa.max() <= 1.0 + 1e-6
    assert np.abs(aa - ref).max() <= 1.0 and (np.abs(aa - ref) > 0).mean() < 0.2            # they differ on the silhouette only
    moved = SAMPLE_POSE.copy(); moved[0, 3] += 0.03
    other = oracle.render_mask(m.vertices, m.faces, mvp_of(K, Hh, Ww, moved), Hh, Ww, anti_aliasing=True)
    iou = lambda a, b: float(np.minimum(a, b).sum() / np.maximum(a, b).sum())
    assert iou(aa, aa) == 1.0 and 0.3 < iou(aa, other) < 0.98
    loss_same, loss_moved = float(((aa - ref) ** 2).sum()), float(((other - ref) ** 2).sum())
    assert loss_same < 0.2 * loss_moved


def _ray_cast_quad(K, H, W, pose, corners):
    """Reference coverage of a planar convex quad by ray casting in float64 (pixel centres, OpenCV camera, z > near)."""
    P = (pose[:3, :3] @ corners.T + pose[:3, 3:4]).T            # corners in the camera frame
    n = np.cross(P[1] - P[0], P[2] - P[0])
    v, u = np.meshgrid(np.arange(H) + 0.5, np.arange(W) + 0.5, indexing="ij")
    d = np.stack([(u - K[0, 2]) / K[0, 0], (v - K[1, 2]) / K[1, 1], np.ones_like(u)], -1)     # ray directions, z = 1
    denom = d @ n
    t = (P[0] @ n) / np.where(np.abs(denom) < 1e-30, 1e-30, denom)
    X = d * t[..., None]
    inside = np.ones((H, W), bool)
    for i in range(4):
        e = P[(i + 1) % 4] - P[i]
        inside &= np.einsum("hwk,k->hw", np.cross(np.broadcast_to(e, X.shape), X - P[i]), n) >= 0
    near, far = 0.001, 10.0
    return inside & (t > near) & (t < far)


def test_triangles_crossing_the_near_plane_are_clipped_and_drawn():
    """A ground plane that runs from behind the camera to 8 m ahead: both of its triangles cross the near plane (and one
    vertex pair lies far outside the guard band).  The clipped render equals ray casting up to boundary pixels, and the
    clipper is reported for both triangles."""
    H, W = 120, 160
    K = scaled_K(H, W)
    corners = np.array([[-3.0, 0.4, -2.0], [3.0, 0.4, -2.0], [3.0, 0.4, 8.0], [-3.0, 0.4, 8.0]])   # y = 0.4 m below the optical axis
    f = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    pose = np.eye(4)
    mvp = mvp_of(K, H, W, pose)
    mask, st = oracle.render_mask(corners.astype(np.float32), f, mvp, H, W, anti_aliasing=False, save=True)
    want = _ray_cast_quad(K.astype(np.float64), H, W, pose, corners)
    assert st[3] == 2                                   # both triangles went through the clipper
    assert want.mean() > 0.3 and mask.mean() > 0.3
    assert (mask != want).mean() < 0.01, (mask != want).sum()
    # antialiased render and its backward run on clipped triangles too (finite values)
    aa, st2 = oracle.render_mask(corners.astype(np.float32), f, mvp, H, W, anti_aliasing=True, save=True)
    assert np.isfinite(aa).all() and abs(float(aa.sum()) - float(mask.sum())) < 0.02 * mask.sum()
    gpos, gmvp = oracle.render_mask_bwd(corners.astype(np.float32), f, mvp, H, W, st2, np.ones((H, W), np.float32))
    assert np.isfinite(gmvp).all()


def test_clipped_grid_has_no_cracks():
    """The same plane as a 12 x 12 grid of triangles: neighbours are cut at bit-identical points, so the clipped render
    has no holes where they meet (every pixel the two-triangle version covers is covered)."""
    H, W = 120, 160
    K = scaled_K(H, W)
    n = 12
    xs, zs = np.linspace(-3, 3, n + 1), np.linspace(-2, 8, n + 1)
    v = np.array([[x, 0.4, z] for z in zs for x in xs], np.float32)
    f = []
    for j in range(n):
        for i in range(n):
            a = j * (n + 1) + i
            f += [[a, a + 1, a + n + 2], [a, a + n + 2, a + n + 1]]
    f = np.array(f, np.int32)
    mvp = mvp_of(K, H, W, np.eye(4))
    grid = oracle.render_mask(v, f, mvp, H, W, anti_aliasing=False)
    two = oracle.render_mask(np.array([[-3, 0.4, -2], [3, 0.4, -2], [3, 0.4, 8], [-3, 0.4, 8]], np.float32),
                             np.array([[0, 1, 2], [0, 2, 3]], np.int32), mvp, H, W, anti_aliasing=False)
    assert two.mean() > 0.3
    # interior of the two-triangle silhouette (one pixel away from its boundary, where the piecewise edge of the grid may
    # legitimately snap differently): no holes
    inner = two.copy()
    inner[1:] &= two[:-1]; inner[:-1] &= two[1:]; inner[:, 1:] &= two[:, :-1]; inner[:, :-1] &= two[:, 1:]
    assert (grid != two).mean() < 0.002 and not (inner & ~grid).any()
