"""CPU model of the rasterizer's lane utilisation on the bench workload (no GPU needed): how many warp iterations the
row-span and the cooperative-fill loops take under the current batching (32 faces per warp, groups of 32 rows, 64x32
units for triangles whose clipped bbox exceeds 96 samples) and how full their lanes are.  Exact row spans are computed
with integer edge functions like the kernels do.  usage: python tools/lane_model.py [views]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402


def tri_rows(X, Y, area, pxlo, pxhi, pylo, pyhi, W, H):
    """covered samples per row of one triangle (exact, top-left rule ignored: ties are measure-zero for statistics)"""
    xs, ys = X.copy(), Y.copy()
    if area < 0:
        xs[[1, 2]] = xs[[2, 1]]; ys[[1, 2]] = ys[[2, 1]]
    bx, by = 8 * W - 8, 8 * H - 8
    px = np.arange(pxlo, pxhi + 1) * 16 - bx
    py = np.arange(pylo, pyhi + 1) * 16 - by
    ok = np.ones((len(py), len(px)), bool)
    for k in range(3):
        k1 = (k + 1) % 3
        ex, ey = xs[k1] - xs[k], ys[k1] - ys[k]
        E = ex * (py[:, None] - ys[k]) - ey * (px[None, :] - xs[k])
        ok &= E >= 0
    return ok.sum(1), ok


def main():
    nviews = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    wl = bench.WORKLOAD
    sets = bench.build_sets(wl, 0, 1)
    sc, mvp = sets[0]["scene"], sets[0]["mvp"]
    H, W = wl["H"], wl["W"]
    span_iters = fill_iters = fill_lanes = span_lanes = 0
    unit_span_iters = unit_fill_iters = unit_fill_lanes = unit_span_lanes = 0
    ntri = ndraw = nbig = 0
    for b in range(min(nviews, mvp.shape[0])):
        for l, m in enumerate(sc["meshes"]):
            v = np.c_[m.vertices, np.ones(len(m.vertices))].astype(np.float32)
            c = v @ mvp[b, l].T.astype(np.float32)
            x = np.rint(c[:, 0] / c[:, 3] * W * 8).astype(np.int64)
            y = np.rint(c[:, 1] / c[:, 3] * H * 8).astype(np.int64)
            f = m.faces
            X, Y = x[f], y[f]
            area = (X[:, 1] - X[:, 0]) * (Y[:, 2] - Y[:, 0]) - (Y[:, 1] - Y[:, 0]) * (X[:, 2] - X[:, 0])
            bx, by = 8 * W - 8, 8 * H - 8
            pxlo = np.maximum((X.min(1) + bx + 15) >> 4, 0); pxhi = np.minimum((X.max(1) + bx) >> 4, W - 1)
            pylo = np.maximum((Y.min(1) + by + 15) >> 4, 0); pyhi = np.minimum((Y.max(1) + by) >> 4, H - 1)
            ok = (area != 0) & (pxlo <= pxhi) & (pylo <= pyhi)
            ntri += len(f)
            for b0 in range(0, len(f), 32):
                rows = []   # covered samples of every row of the batch's small triangles, in triangle order
                for t in range(b0, min(b0 + 32, len(f))):
                    if not ok[t]:
                        continue
                    ndraw += 1
                    w_, h_ = pxhi[t] - pxlo[t] + 1, pyhi[t] - pylo[t] + 1
                    cnt, cov = tri_rows(X[t], Y[t], area[t], pxlo[t], pxhi[t], pylo[t], pyhi[t], W, H)
                    if w_ * h_ > 96:
                        nbig += 1
                        for uy in range(0, h_, 32):
                            for ux in range(0, w_, 64):
                                sub = cov[uy:uy + 32, ux:ux + 64]
                                if not sub.any():
                                    continue   # (most empty windows are voided at creation)
                                unit_span_iters += 1; unit_span_lanes += sub.shape[0]
                                T = int(sub.sum())
                                unit_fill_iters += -(-T // 32); unit_fill_lanes += T
                    else:
                        rows.extend(cnt.tolist())
                for g in range(0, len(rows), 32):
                    grp = rows[g:g + 32]
                    span_iters += 1; span_lanes += len(grp)
                    T = int(sum(grp))
                    fill_iters += -(-T // 32); fill_lanes += T
    pr = lambda name, it, lanes: print("%-34s %8d warp iterations, lanes %5.1f %% full" % (name, it, 100.0 * lanes / max(32 * it, 1)))
    print("%d views: %d triangles, %d drawable, %d deferred" % (min(nviews, mvp.shape[0]), ntri, ndraw, nbig))
    pr("small: row-span groups", span_iters, span_lanes)
    pr("small: fill (32 samples)", fill_iters, fill_lanes)
    pr("deferred units: row-span", unit_span_iters, unit_span_lanes)
    pr("deferred units: fill", unit_fill_iters, unit_fill_lanes)
    tot_it = span_iters + fill_iters + unit_span_iters + unit_fill_iters
    ideal = -(-(span_lanes + unit_span_lanes) // 32) + -(-(fill_lanes + unit_fill_lanes) // 32)
    print("total %d iterations; perfectly packed lanes would need %d (%.2fx fewer)" % (tot_it, ideal, tot_it / max(ideal, 1)))


if __name__ == "__main__":
    main()
