"""Developer tool: per-phase cycle counters of ehb_k_tiles (EHB_STATS build) on the bench scene.
   EHB_LIB=easyhec_b200/libehb_stats.so python tools/tile_stats.py [headline|inview] [H W]"""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
from easyhec_b200._lib import Context  # noqa: E402

wl = dict(b.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "headline"])
if len(sys.argv) > 3:
    wl["H"], wl["W"] = int(sys.argv[2]), int(sys.argv[3])
H, W, B = wl["H"], wl["W"], wl["B"]
s = b.build_sets(wl, 0, 1)[0]
ctx = Context("cuda:0")
ctx.set_pipelines(1)
ids = [ctx.register_mesh(m.vertices, m.faces) for m in s["scene"]["meshes"]]
ref = ctx.register_ref(ctx.render_binary_batch(ids, torch.from_numpy(s["mvp_gt"]).cuda(), H, W))
mvp = torch.from_numpy(s["mvp"]).cuda()
for _ in range(3):
    ctx.render_views_fused(ids, mvp, ref, H, W, backward=True)
ctx.debug_counters(reset=True)
ctx.debug_buffer(0)
n = 10
for _ in range(n):
    ctx.render_views_fused(ids, mvp, ref, H, W, backward=True)
c = ctx.debug_counters()
names = ["tiles", "links", "tiles_nl0", "pairs", "multi_round", "A_windows", "B_pairs", "C_weights", "D_masks", "E_compose",
         "F_backward", "whole_tile"]
t = max(c[0], 1)
print({k: v / n for k, v in zip(names[:5], c[:5])})
print({k: round(v / t) for k, v in zip(names[5:], c[5:12])}, "(cycles per tile, thread 0)")

import numpy as np
buf = np.array(ctx.debug_buffer(4096 * 8), dtype=np.uint64).reshape(4096, 8).astype(np.int64)
ok = buf[:, 6] > 0
rows = buf[ok]
order = np.argsort(-rows[:, 6])
print("slowest tiles of the last pass (cycles, thread 0): A B C D E F whole | links pairs")
for r in rows[order[:12]]:
    print("  %6d %6d %6d %6d %6d %6d %7d | %d %d" % (r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7] & 255, r[7] >> 8))
nl = rows[:, 7] & 255
for k in range(0, 8):
    m = nl == k
    if m.any():
        print("  tiles with %d links: %4d, mean whole %.0f cycles, max %d" % (k, m.sum(), rows[m, 6].mean(), rows[m, 6].max()))
