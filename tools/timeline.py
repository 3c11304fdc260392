"""Developer tool: timeline of one fused pass (EHB_TIMELINE build) on the bench scene -- start / end of every warp or CTA
of the five kernels on the global timer: per-kernel span, distribution of the entries' durations, how many SMs are
still busy towards the end (the tail), gaps between the kernels.
   make -C easyhec_b200/csrc EXTRA=-DEHB_TIMELINE OUT=../libehb_tl.so
   EHB_LIB=easyhec_b200/libehb_tl.so python tools/timeline.py [headline|inview] [items] [H W]"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
from easyhec_b200._lib import Context  # noqa: E402

wl = dict(b.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "headline"])
items = int(sys.argv[2]) if len(sys.argv) > 2 else wl["B"]
if len(sys.argv) > 4:
    wl["H"], wl["W"] = int(sys.argv[3]), int(sys.argv[4])
H, W = wl["H"], wl["W"]
s = b.build_sets(wl, 0, 1)[0]
ctx = Context("cuda:0")
ctx.set_pipelines(1)
ids = [ctx.register_mesh(m.vertices, m.faces) for m in s["scene"]["meshes"]]
ref = ctx.register_ref(ctx.render_binary_batch(ids, torch.from_numpy(s["mvp_gt"]).cuda(), H, W))
mvp = torch.from_numpy(s["mvp"]).cuda()[:items].contiguous()
ref = ref[:items]
masks = torch.empty((items, H, W), dtype=torch.float32, device="cuda")
loss = torch.empty((items,), dtype=torch.float64, device="cuda")
g = torch.empty((items, len(ids), 4, 4), dtype=torch.float64, device="cuda")
N = 16384
ctx.debug_buffer(0)          # allocates the buffer
for _ in range(5):
    ctx.render_views_fused(ids, mvp, ref, H, W, backward=True, out=(masks, loss, g))
torch.cuda.synchronize()
buf = np.array(ctx.debug_buffer(6 * N * 2), dtype=np.uint64).reshape(6, N, 2)
names = ["table", "front", "raster(warp)", "raster_big(warp)", "tiles(cta)", "raster-stream(warp)"]
t_first = None
spans = {}
for k in range(6):
    st = buf[k, :, 0].astype(np.int64)
    en = (buf[k, :, 1] >> np.uint64(8)).astype(np.int64) & ((1 << 56) - 1)
    st = st & ((1 << 56) - 1)
    sm = (buf[k, :, 1] & np.uint64(255)).astype(np.int64)
    ok = buf[k, :, 0] != 0
    if not ok.any():
        continue
    st, en, sm = st[ok], en[ok], sm[ok]
    if t_first is None:
        t_first = st.min()
    d = (en - st) / 1e3
    k0, k1 = st.min(), en.max()
    spans[k] = (k0, k1)
    rel_end = (k1 - en) / 1e3
    print("%-20s entries %5d  start %+7.2f us  span %6.2f us | entry us: mean %.2f p50 %.2f p90 %.2f p99 %.2f max %.2f | first start spread %.2f us" % (
        names[k], ok.sum(), (k0 - t_first) / 1e3, (k1 - k0) / 1e3, d.mean(), np.percentile(d, 50), np.percentile(d, 90),
        np.percentile(d, 99), d.max(), (st.max() - k0) / 1e3))
    # SMs still busy t us before the end of the kernel
    line = []
    for back in (0.5, 1, 2, 3, 5, 8, 12):
        t = k1 - int(back * 1e3)
        busy = len(set(sm[(st <= t) & (en > t)]))
        act = int(((st <= t) & (en > t)).sum())
        line.append("-%gus: %d SMs / %d" % (back, busy, act))
    print("    busy before the end:  " + "   ".join(line))
    # busy time per SM as a fraction of the span (entries may overlap on an SM: union of intervals)
    fr = []
    for s_ in range(int(sm.max()) + 1):
        iv = sorted(zip(st[sm == s_], en[sm == s_]))
        tot, cur0, cur1 = 0, None, None
        for a, e in iv:
            if cur1 is None or a > cur1:
                if cur1 is not None:
                    tot += cur1 - cur0
                cur0, cur1 = a, e
            else:
                cur1 = max(cur1, e)
        if cur1 is not None:
            tot += cur1 - cur0
        fr.append(tot / max(k1 - k0, 1))
    print("    SM busy fraction of the span: mean %.2f min %.2f max %.2f" % (np.mean(fr), np.min(fr), np.max(fr)))
order = sorted(spans, key=lambda k: spans[k][0])
for a, c in zip(order[:-1], order[1:]):
    print("gap %s end -> %s first start: %.2f us" % (names[a], names[c], (spans[c][0] - spans[a][1]) / 1e3))
print("whole pass: %.2f us" % ((max(v[1] for v in spans.values()) - min(v[0] for v in spans.values())) / 1e3))
