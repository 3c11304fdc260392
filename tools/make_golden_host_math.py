"""Generate tests/golden/host_math.npz by RUNNING the reference's own Python for the host-side math.

Run in the build container only (/root/reference is not on the GPU box).  The reference modules are
imported from /root/reference with import shims for packages that are absent here and that the
functions under test do not exercise numerically:
  * pytorch3d.transforms.so3.hat  -> the 3-line skew-symmetric definition (pytorch3d is not installed;
    this is the one piece restated rather than executed -- se3 goldens are "pinned modulo hat");
  * loguru / multipledispatch / pn_utils -> import-time only;
  * Tensor.cuda() -> identity (the reference hard-codes .cuda(); arithmetic is identical on CPU fp32).
Pinned functions: K_to_projection, transform_pos (easyhec/utils/nvdiffrast_utils.py:5-18),
se3_exp_map (easyhec/utils/pytorch3d_se3.py:46-130), se3_log_map(backend='opencv')
(easyhec/utils/utils_3d.py:308-335), the projection/pose chain of render_mask
(easyhec/structures/nvdiffrast_renderer.py:33-37) and RBSolver's init dof (rb_solver.py:31-33).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("EHB_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "host_math.npz")


def _shim():
    def hat(v):
        x, y, z = v.unbind(1)
        o = torch.zeros_like(x)
        return torch.stack([o, -z, y, z, o, -x, -y, x, o], 1).reshape(-1, 3, 3)

    p3d = types.ModuleType("pytorch3d")
    tr = types.ModuleType("pytorch3d.transforms")
    so3 = types.ModuleType("pytorch3d.transforms.so3")
    se3 = types.ModuleType("pytorch3d.transforms.se3")
    so3.hat = hat
    so3.so3_log_map = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError())
    tr.so3, tr.se3 = so3, se3
    p3d.transforms = tr
    sys.modules.update({"pytorch3d": p3d, "pytorch3d.transforms": tr, "pytorch3d.transforms.so3": so3,
                        "pytorch3d.transforms.se3": se3})
    lg = types.ModuleType("loguru")
    lg.logger = types.SimpleNamespace(warning=lambda *a, **k: None, info=lambda *a, **k: None)
    sys.modules["loguru"] = lg
    md = types.ModuleType("multipledispatch")
    md.dispatch = lambda *a, **k: (lambda f: f)
    sys.modules["multipledispatch"] = md
    torch.Tensor.cuda = lambda self, *a, **k: self


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    _shim()
    pkg = types.ModuleType("easyhec"); pkg.__path__ = [os.path.join(REF, "easyhec")]
    up = types.ModuleType("easyhec.utils"); up.__path__ = [os.path.join(REF, "easyhec", "utils")]
    sys.modules["easyhec"], sys.modules["easyhec.utils"] = pkg, up
    pn = types.ModuleType("easyhec.utils.pn_utils")
    pn.to_array = lambda x, dtype=float: np.asarray(x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x)
    sys.modules["easyhec.utils.pn_utils"] = pn
    nu = _load("easyhec.utils.nvdiffrast_utils", "easyhec/utils/nvdiffrast_utils.py")
    _load("easyhec.utils.pytorch3d_se3", "easyhec/utils/pytorch3d_se3.py")
    u3 = _load("easyhec.utils.utils_3d", "easyhec/utils/utils_3d.py")

    rng = np.random.RandomState(0)
    out = {}
    # K_to_projection on the reference's default intrinsics and two scaled ones
    Ks, HWs, projs = [], [], []
    for (fx, fy, cx, cy, W, H) in [(906.805, 906.680, 650.198, 367.714, 1280, 720),
                                   (1352.21, 1352.43, 963.35, 529.40, 1920, 1080),
                                   (386.32, 385.39, 331.31, 239.81, 640, 480), (90.68, 90.67, 64.0, 64.0, 128, 128)]:
        K = torch.tensor([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=torch.float32)
        Ks.append(K.numpy()); HWs.append([H, W]); projs.append(nu.K_to_projection(K, H, W).numpy())
    out["K"], out["HW"], out["proj"] = np.stack(Ks), np.array(HWs), np.stack(projs)
    # se3_exp_map / se3_log_map(opencv)
    dof = torch.from_numpy(np.concatenate([rng.uniform(-1, 1, (16, 3)), rng.uniform(-2.5, 2.5, (16, 3))], 1)).float()
    dof[0, 3:] = 0.0           # theta -> 0 (eps clamp)
    dof[1, 3:] = 1e-3
    T = u3.se3_exp_map(dof)
    out["dof"], out["exp"] = dof.numpy(), T.numpy()
    out["log"] = u3.se3_log_map(T, backend="opencv").numpy()
    # RBSolver init: the Franka offline example's init_Tc_c2b (configs/franka/example_franka_offline.yaml:5-8)
    import yaml
    y = yaml.safe_load(open(os.path.join(REF, "configs/franka/example_franka_offline.yaml")))
    init = np.array(y["model"]["rbsolver"]["init_Tc_c2b"], dtype=np.float32)
    out["init_Tc_c2b"] = init
    out["init_dof"] = u3.se3_log_map(torch.as_tensor(init)[None].permute(0, 2, 1), eps=1e-5, backend="opencv")[0].numpy()
    # render_mask's clip-space chain on a few random points
    verts = torch.from_numpy(rng.uniform(-0.3, 0.3, (64, 3))).float()
    pose = torch.from_numpy(np.array([[0.99638397, -0.0846324, 0.00750877, -0.20668708],
                                      [-0.00875172, -0.19013488, -0.9817189, 0.08405855],
                                      [0.0845129, 0.97810328, -0.19018805, 0.77892876],
                                      [0., 0., 0., 1.]], dtype=np.float32))
    blender2opencv = torch.tensor([[1, 0, 0, 0], [0, -1, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]]).float()
    opencv2blender = torch.inverse(blender2opencv)
    proj = nu.K_to_projection(torch.from_numpy(Ks[0]), 720, 1280)
    mvp = proj @ (opencv2blender @ pose)
    out["chain_verts"], out["chain_pose"], out["chain_mvp"] = verts.numpy(), pose.numpy(), mvp.numpy()
    out["chain_clip"] = nu.transform_pos(mvp, verts)[0].numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
