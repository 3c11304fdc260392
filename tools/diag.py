"""Developer diagnostics on the bench workload (run on the GPU box):
  * host enqueue time per step vs device time per step (is the eager loop launch-bound?)
  * with an -DEHB_TIMING build (EHB_LIB=...): k_tiles phase cycle counters, tiles / links / pairs per step
usage: [EHB_LIB=easyhec_b200/libehb_timing.so] [EHB_PIPES=n] python tools/diag.py [steps]"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from easyhec_b200._lib import Context  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
wl = bench.WORKLOAD
B, H, W, R = wl["B"], wl["H"], wl["W"], wl["ring"]
dev = torch.device("cuda", 0)
ctx = Context(dev)
if os.environ.get("EHB_PIPES"):
    ctx.set_pipelines(int(os.environ["EHB_PIPES"]))
sets = bench.build_sets(wl, 0, R)
meshes = sets[0]["scene"]["meshes"]
ids = [ctx.register_mesh(m.vertices, m.faces) for m in meshes]
L = len(ids)
mvp_dev = [torch.from_numpy(s["mvp"]).to(dev) for s in sets]
ref_dev = [ctx.render_binary_batch(ids, torch.from_numpy(s["mvp_gt"]).to(dev), H, W).to(torch.float32) for s in sets]
masks = [torch.empty((B, H, W), dtype=torch.float32, device=dev) for _ in range(R)]
loss = torch.empty((B,), dtype=torch.float64, device=dev)
gmvp = torch.empty((B, L, 4, 4), dtype=torch.float64, device=dev)


def step(k):
    s = k % R
    ctx.render_views_fused(ids, mvp_dev[s], ref_dev[s], H, W, backward=True, out=(masks[s], loss, gmvp))


for k in range(20):
    step(k)
torch.cuda.synchronize()
out = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for k in range(steps):
    step(k)
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
out["host_enqueue_us_per_step"] = 1e6 * (t1 - t0) / steps
out["device_us_per_step"] = 1e3 * e0.elapsed_time(e1) / steps
# the same step replayed from a CUDA graph (no host launch cost)
try:
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        step(0)
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            for k in range(R):
                step(k)
    torch.cuda.synchronize()
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    e0.record()
    n = max(steps // R, 1)
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    out["graph_us_per_step"] = 1e3 * e0.elapsed_time(e1) / (n * R)
except Exception as e:  # noqa: BLE001
    out["graph_error"] = str(e)[:200]
# phase counters (EHB_TIMING builds only; zeros otherwise)
ctx.set_pipelines(1)
ctx.debug_counters(reset=True)
nst = 20
for k in range(nst):
    step(k)
c = [int(x) for x in ctx.debug_counters(reset=True)]
if any(c):
    names = ["setup", "window", "pairlist", "weights", "gather", "compose", "backward", "-", "-", "queue"]
    tot = sum(c[:10]) or 1
    out["tiles_phase_pct"] = {names[i]: round(100.0 * c[i] / tot, 1) for i in range(10) if names[i] != "-"}
    out["tiles_per_step"] = c[10] / nst
    out["links_per_tile"] = c[11] / max(c[10], 1)
    out["pairs_per_tile"] = c[12] / max(c[10], 1)
    out["cycles_per_tile_thread0"] = tot / max(c[10], 1)
    out["worst_tile_cycles"] = c[13]
    out["tiles_over_50k"] = c[14] / nst
    out["tiles_over_100k"] = c[15] / nst
print(json.dumps(out))
