#!/usr/bin/env python
"""Dump golden vectors from the REAL reference stack (EasyHeC + nvdiffrast) into tests/golden/nvdiffrast_golden.npz.

The rasterize / antialias arithmetic of the render_mask path lives in nvdiffrast, which is neither vendored in the
reference tree nor installable in the build container, so this repo's oracle is "parity unpinned" for it (DESIGN.md
section 2).  This script is the route to pinning: run it ONCE on any machine that has a CUDA GPU, nvdiffrast and a
checkout of ootts/EasyHeC, commit the file it writes, and tests/test_nvdiffrast_golden.py will (a) select the fill rule
(`ehb_ctx_set_fill_rule`, `oracle.FILL_RULE`) under which the binary masks agree bit for bit, (b) hold the oracle and
the CUDA kernels to the recorded masks and gradients.

    python tools/dump_nvdiffrast_golden.py --reference /path/to/EasyHeC [--out tests/golden/nvdiffrast_golden.npz]

What is run is exactly the reference operator: easyhec/structures/nvdiffrast_renderer.py:25-48 (`render_mask`, with and
without antialiasing) and :50-73 (`batch_render_mask`), called like rb_solver.py:60-67 does, on inputs taken from this
repo's committed fixtures so that both sides see identical (verts, faces, K, pose) arrays:
  case "zero128"   all 8 xArm7 links at qpos 0 packed into one mesh (= assets/xarm7_zeropos.ply), 128x128, sample pose
  case "link480"   link 6 alone, 640x480, sample pose
  case "squares"   axis-aligned squares at integer / half-integer / quarter pixel offsets, 64x64 (the tie-rule probe)
  case "bench0"    view 0 of bench.py's headline scene, every link on its own, 1280x720 (+ the clamp-sum composition)
Each case stores inputs, the binary mask, the antialiased mask, and d(sum(mask * dy))/d(object_pose) for a fixed dy.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def cases():
    from easyhec_b200.meshio import Mesh, concat_meshes
    from easyhec_b200.scenes import SAMPLE_POSE, load_xarm7, scaled_K
    import importlib.util
    fx = load_xarm7()
    out = []
    m = concat_meshes([mm.transformed(T) for mm, T in zip(fx["meshes"], fx["fk_zero"])])
    out.append(("zero128", m.vertices, m.faces, scaled_K(128, 128), SAMPLE_POSE, 128, 128))
    m = fx["meshes"][6].transformed(fx["fk_zero"][6])
    out.append(("link480", m.vertices, m.faces, scaled_K(480, 640), SAMPLE_POSE, 480, 640))
    K = np.array([[64.0, 0, 32.0], [0, 64.0, 32.0], [0, 0, 1]], np.float32)
    for i, (off, size) in enumerate([(0.0, 8), (0.5, 8), (0.25, 4), (0.5, 1)]):
        a, b = (10 + off - 32) / 64.0, (10 + off + size - 32) / 64.0
        v = np.array([[a, a, 1], [b, a, 1], [b, b, 1], [a, b, 1]], np.float32)
        f = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
        out.append(("squares%d" % i, v, f, K, np.eye(4), 64, 64))
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    s = b.build_sets(b.WORKLOAD, 0, 1)[0]
    sc = s["scene"]
    for l, mm in enumerate(sc["meshes"]):
        pose = s["Tc"] @ sc["link_poses"][0, l].astype(np.float64)
        out.append(("bench0_link%d" % l, mm.vertices, mm.faces, sc["K"], pose, 720, 1280))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="checkout of ootts/EasyHeC (its easyhec/ package is imported)")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "nvdiffrast_golden.npz"))
    args = ap.parse_args()
    sys.path.insert(0, args.reference)
    import torch
    import nvdiffrast
    from easyhec.structures.nvdiffrast_renderer import NVDiffrastRenderer   # the reference's own operator

    arrs = {"nvdiffrast_version": np.array(getattr(nvdiffrast, "__version__", "unknown")),
            "torch_version": np.array(torch.__version__), "gpu": np.array(torch.cuda.get_device_name(0))}
    names = []
    for name, v, f, K, pose, H, W in cases():
        r = NVDiffrastRenderer([H, W])
        vt = torch.from_numpy(np.ascontiguousarray(v, np.float32)).cuda()
        ft = torch.from_numpy(np.ascontiguousarray(f, np.int32)).cuda()
        Kt = torch.from_numpy(np.asarray(K, np.float32)).cuda()
        pt = torch.tensor(np.asarray(pose), dtype=torch.float32, device="cuda", requires_grad=True)
        binary = r.render_mask(vt, ft, Kt, pt.detach(), anti_aliasing=False)
        aa = r.render_mask(vt, ft, Kt, pt, anti_aliasing=True)
        dy = torch.from_numpy(np.random.RandomState(1).randn(H, W).astype(np.float32)).cuda()
        (aa * dy).sum().backward()
        names.append(name)
        arrs.update({name + "_verts": v.astype(np.float32), name + "_faces": f.astype(np.int32),
                     name + "_K": np.asarray(K, np.float32), name + "_pose": np.asarray(pose, np.float32),
                     name + "_HW": np.array([H, W]), name + "_binary": np.packbits(binary.cpu().numpy().astype(bool), axis=-1),
                     name + "_aa": aa.detach().cpu().numpy(), name + "_g_pose": pt.grad.cpu().numpy()})
        print("%-14s covered %6d px  aa sum %.3f  |g_pose| %.3e" % (name, int(binary.sum()), float(aa.sum()), float(pt.grad.abs().max())))
    arrs["cases"] = np.array(names)
    np.savez_compressed(args.out, **arrs)
    print("wrote", args.out, os.path.getsize(args.out), "bytes -- commit it; tests/test_nvdiffrast_golden.py consumes it")


if __name__ == "__main__":
    main()
