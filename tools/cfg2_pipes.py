"""Developer tool: cfg2 (10 views 640x480 pose solve) iterations/s for 1..4 pipelines."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from easyhec_b200._lib import Context
from easyhec_b200.scenes import make_scene, perturb_pose
from easyhec_b200.solver import PoseSolver
from util import scene_mvps
B, H, W = 10, 480, 640
sc = make_scene(B, H, W, links="xarm7", seed=0)
for pipes in (1, 2, 3, 4):
    ctx = Context("cuda:0"); ctx.set_pipelines(pipes)
    ids = [ctx.register_mesh(m.vertices, m.faces) for m in sc["meshes"]]
    ref = ctx.render_binary_batch(ids, torch.from_numpy(scene_mvps(sc, H, W)).cuda(), H, W)
    init = perturb_pose(sc["Tc_c2b"], np.random.RandomState(0), 0.03, 3.0)
    s = PoseSolver(sc["meshes"], sc["link_poses"], sc["K"], ref, init, H, W, ctx=ctx)
    s.step(2); torch.cuda.synchronize()
    t0 = time.perf_counter(); s.step(512); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("pipes %d: %.0f iterations/s (%.1f us), err %s" % (pipes, 512 / dt, 1e6 * dt / 512, s.pose_error(sc["Tc_c2b"])), flush=True)
    ctx.close()
