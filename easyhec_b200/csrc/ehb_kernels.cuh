// ehb_kernels.cuh -- the four kernels of one rasterizer pass (count -> alloc -> fill -> raster).
//
// Work decomposition (B200: 148 SMs, 227 KB smem/SM, 126 MB L2):
//   * the screen of every item (a camera view, or one render of a batch) is cut into 32x32-pixel tiles;
//     a tile's CTA keeps one 64-bit (depth key | triangle id) plane PER LINK in shared memory, because the
//     reference antialiases every link separately before summing (rb_solver.py:62-68);
//   * k_count  : one thread per (item, triangle): transform, snap, cull, bbox -> per-(tile, link) counts;
//   * k_alloc  : one warp per tile: contiguous, 16-byte aligned segment of the pair buffer per tile, list of
//                non-empty tiles; EMPTY tiles are finished right here as a pure float4 stream
//                (mask = 0, loss += ref^2), which is the HBM-bound part of the frame;
//   * k_fill   : one thread per (item, triangle): scatter triangle ids into the tile segments;
//   * k_raster : persistent CTAs pull non-empty tiles from a queue: coverage + nearest depth with shared
//                memory atomicMin, antialias forward of every link, sum / clamp / loss / dL/dmask, antialias
//                backward reduced by warp shuffles straight to d loss / d mvp[item, link] (fp64 atomics).
// No intermediate image (rast, colour, work queue, clip-space vertex buffer) ever goes to HBM.
#pragma once
#include "ehb_device.cuh"

#define EHB_T 32            // tile interior
#define EHB_RS 35           // plane row stride = T + max halo (1 low, 2 high)
#define EHB_NP (EHB_RS * EHB_RS)

enum { EHB_MODE_FUSED = 0, EHB_MODE_AA_FWD = 1, EHB_MODE_AA_BWD = 2, EHB_MODE_UNION = 3, EHB_MODE_UNION_VAR = 4 };

struct EhbPairEnt;
struct EhbCounters {
    unsigned long long pairCursor;
    unsigned long long nNeedClip;
    unsigned int nHeavy;     // non-empty tiles with many triangles: listed from the front of tileList
    unsigned int nLight;     // the others: listed from the back; the raster queue serves the heavy ones first
    unsigned int workCursor;
    unsigned int flags;
};
#define EHB_HEAVY_PAIRS 192

struct EhbParams {
    int H, W, ntx, nty, ntiles;
    int items, L, Lk, Ftot;
    int hlo, hhi;
    int mode, rule, do_bwd, clamp;
    float invB;
    const float* mvp;     // [items, L, 16]
    uint32_t* range;      // [items, Ftot] packed tile range of every triangle
    uint32_t* cnt;        // [items * ntiles * Lk]
    uint32_t* start;
    uint32_t* cur;
    uint32_t* tileList;   // [items * ntiles]
    uint32_t* pairs;
    unsigned long long pairCap;
    EhbCounters* ctr;
    const float* ref;     // [items, H, W]  FUSED
    const uint8_t* ref_u8;// same, as bytes (either ref or ref_u8)
    float* masks;         // [items, H, W]  FUSED / AA_FWD
    double* loss;         // [items]
    double* gmvp;         // [items, L, 16]
    float* gpos;          // [V, 4] (AA_BWD, single link) or NULL
    const float* dy;      // [items, H, W]  AA_BWD
    uint8_t* out_u8;      // [items, H, W]  UNION
    float* score;         // [items / C]    UNION_VAR
    int C;
    struct EhbPairEnt* pairSpill;   // [gridDim.x of k_raster][spillCap] overflow of the shared-memory pair lists
    int spillCap;
};

#define EHB_RANGE_NONE 1u  // lo_x = 1 > hi_x = 0

__device__ __forceinline__ int ehb_find_link(const EhbRobot& rb, int g)
{
    int l = 0;
    while (l + 1 < rb.L && g >= rb.foff[l + 1]) l++;
    return l;
}

__device__ __forceinline__ void ehb_load_mvp(const float* __restrict__ src, float* m)
{
    const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float4 v = __ldg(s4 + i);
        m[4 * i] = v.x; m[4 * i + 1] = v.y; m[4 * i + 2] = v.z; m[4 * i + 3] = v.w;
    }
}

// ------------------------------------------------------------------------------------------------ k_count
__global__ void __launch_bounds__(256) ehb_k_count(const __grid_constant__ EhbRobot rb,
                                                   const __grid_constant__ EhbParams p)
{
    const int item = blockIdx.y;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x == 0) {
        if (item == 0 && threadIdx.x == 0) {
            p.ctr->pairCursor = 0ull;
            p.ctr->nHeavy = 0u;
            p.ctr->nLight = 0u;
            p.ctr->workCursor = 0u;
        }
        if (p.loss && threadIdx.x == 0) p.loss[item] = 0.0;
        if (p.gmvp)
            for (int i = threadIdx.x; i < p.L * 16; i += blockDim.x) p.gmvp[(size_t)item * p.L * 16 + i] = 0.0;
        if (p.score && p.C > 0 && item % p.C == 0 && threadIdx.x == 0) p.score[item / p.C] = 0.f;
    }
    if (g >= p.Ftot) return;
    const int l = ehb_find_link(rb, g);
    float m[16];
    ehb_load_mvp(p.mvp + ((size_t)item * p.L + l) * 16, m);
    EhbTri s;
    const int st = ehb_tri_setup(rb.link[l], m, g - rb.foff[l], p.H, p.W, s);
    uint32_t packed = EHB_RANGE_NONE;
    if (st == 2) {
        atomicAdd(&p.ctr->nNeedClip, 1ull);
        atomicOr(&p.ctr->flags, 2u);
    } else if (st == 0) {
        const int txlo = max(0, (s.pxlo - p.hhi) >> 5), txhi = min(p.ntx - 1, (s.pxhi + p.hlo) >> 5);
        const int tylo = max(0, (s.pylo - p.hhi) >> 5), tyhi = min(p.nty - 1, (s.pyhi + p.hlo) >> 5);
        packed = (uint32_t)txlo | ((uint32_t)txhi << 8) | ((uint32_t)tylo << 16) | ((uint32_t)tyhi << 24);
        const int lb = p.Lk == 1 ? 0 : l;
        for (int ty = tylo; ty <= tyhi; ty++)
            for (int tx = txlo; tx <= txhi; tx++)
                atomicAdd(&p.cnt[((size_t)item * p.ntiles + ty * p.ntx + tx) * p.Lk + lb], 1u);
    }
    p.range[(size_t)item * p.Ftot + g] = packed;
}

// ------------------------------------------------------------------------------------------------ k_alloc
__device__ __forceinline__ double ehb_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// A tile no triangle touches: mask = 0 and loss += sum ref^2, streamed by one warp.
__device__ __forceinline__ void ehb_stream_empty_tile(const EhbParams& p, int item, int tx, int ty, int lane)
{
    const int x0 = tx * EHB_T, y0 = ty * EHB_T;
    const int H = p.H, W = p.W;
    const size_t ibase = (size_t)item * H * W;
    if (p.mode == EHB_MODE_FUSED || p.mode == EHB_MODE_AA_FWD) {
        double acc = 0.0;
        const bool vec = (W & 3) == 0 && ((((uintptr_t)p.masks) | ((uintptr_t)p.ref)) & 15) == 0;
        if (vec) {
            const int cx = x0 + 4 * (lane & 7);
#pragma unroll
            for (int it = 0; it < 8; it++) {
                const int py = y0 + it * 4 + (lane >> 3);
                if (py < H && cx < W) {
                    const size_t o = ibase + (size_t)(H - 1 - py) * W + cx;
                    if (p.ref) {
                        const float4 r = __ldg(reinterpret_cast<const float4*>(p.ref + o));
                        acc += (double)(r.x * r.x) + (double)(r.y * r.y) + (double)(r.z * r.z) + (double)(r.w * r.w);
                    } else if (p.ref_u8) {
                        const uchar4 r = __ldg(reinterpret_cast<const uchar4*>(p.ref_u8 + o));
                        acc += (double)((r.x != 0) + (r.y != 0) + (r.z != 0) + (r.w != 0));
                    }
                    if (p.masks) *reinterpret_cast<float4*>(p.masks + o) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        } else {
            for (int i = lane; i < EHB_T * EHB_T; i += 32) {
                const int px = x0 + (i & 31), py = y0 + (i >> 5);
                if (px < W && py < H) {
                    const size_t o = ibase + (size_t)(H - 1 - py) * W + px;
                    if (p.ref) { const float r = __ldg(p.ref + o); acc += (double)(r * r); }
                    else if (p.ref_u8) acc += (double)(__ldg(p.ref_u8 + o) != 0);
                    if (p.masks) p.masks[o] = 0.f;
                }
            }
        }
        if (p.loss && (p.ref || p.ref_u8)) {
            acc = ehb_warp_sum(acc);
            if (lane == 0 && acc != 0.0) atomicAdd(&p.loss[item], acc);
        }
    } else if (p.mode == EHB_MODE_UNION) {
        if ((W & 3) == 0 && (((uintptr_t)p.out_u8) & 3) == 0) {
            const int cx = x0 + 4 * (lane & 7);
#pragma unroll
            for (int it = 0; it < 8; it++) {
                const int py = y0 + it * 4 + (lane >> 3);
                if (py < H && cx < W)
                    *reinterpret_cast<uint32_t*>(p.out_u8 + ibase + (size_t)(H - 1 - py) * W + cx) = 0u;
            }
        } else {
            for (int i = lane; i < EHB_T * EHB_T; i += 32) {
                const int px = x0 + (i & 31), py = y0 + (i >> 5);
                if (px < W && py < H) p.out_u8[ibase + (size_t)(H - 1 - py) * W + px] = 0;
            }
        }
    }
}

__global__ void __launch_bounds__(256) ehb_k_alloc(const __grid_constant__ EhbParams p)
{
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= p.items * p.ntiles) return;
    const int item = wid / p.ntiles, tile = wid - item * p.ntiles;
    const size_t bin0 = (size_t)wid * p.Lk;
    const uint32_t c = lane < p.Lk ? p.cnt[bin0 + lane] : 0u;
    uint32_t inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    bool empty = total == 0;
    unsigned long long base = 0;
    if (!empty) {
        const uint32_t padded = (total + 3u) & ~3u;
        if (lane == 0) base = atomicAdd(&p.ctr->pairCursor, (unsigned long long)padded);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base + padded > p.pairCap) {  // pair buffer too small: flag it, the host grows the buffer and reruns
            if (lane == 0) atomicOr(&p.ctr->flags, 1u);
            if (lane < p.Lk) p.cnt[bin0 + lane] = 0u;
            empty = true;
        }
    }
    if (empty) {
        if (lane < p.Lk) { p.start[bin0 + lane] = 0xFFFFFFFFu; p.cur[bin0 + lane] = 0xFFFFFFFFu; }
        ehb_stream_empty_tile(p, item, tile % p.ntx, tile / p.ntx, lane);
        return;
    }
    if (lane < p.Lk) {
        const uint32_t s = (uint32_t)base + (inc - c);
        p.start[bin0 + lane] = s;
        p.cur[bin0 + lane] = s;
    }
    if (lane == 0) {
        if (total >= EHB_HEAVY_PAIRS) p.tileList[atomicAdd(&p.ctr->nHeavy, 1u)] = (uint32_t)wid;
        else p.tileList[(unsigned)(p.items * p.ntiles) - 1u - atomicAdd(&p.ctr->nLight, 1u)] = (uint32_t)wid;
    }
}

// ------------------------------------------------------------------------------------------------ k_fill
__global__ void __launch_bounds__(256) ehb_k_fill(const __grid_constant__ EhbRobot rb,
                                                  const __grid_constant__ EhbParams p)
{
    const int item = blockIdx.y;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.Ftot) return;
    const uint32_t r = p.range[(size_t)item * p.Ftot + g];
    const int txlo = r & 255, txhi = (r >> 8) & 255, tylo = (r >> 16) & 255, tyhi = r >> 24;
    if (txlo > txhi) return;
    const int l = ehb_find_link(rb, g);
    const uint32_t entry = ((uint32_t)l << EHB_LINK_SHIFT) | (uint32_t)(g - rb.foff[l]);
    const int lb = p.Lk == 1 ? 0 : l;
    for (int ty = tylo; ty <= tyhi; ty++)
        for (int tx = txlo; tx <= txhi; tx++) {
            const size_t bin = ((size_t)item * p.ntiles + ty * p.ntx + tx) * p.Lk + lb;
            if (p.start[bin] == 0xFFFFFFFFu) continue;  // tile dropped by an overflowing alloc
            const uint32_t slot = atomicAdd(&p.cur[bin], 1u);
            p.pairs[slot] = entry;
        }
}

// ------------------------------------------------------------------------------------------------ k_raster
//
// One CTA (128 threads) per non-empty tile; persistent CTAs pull tiles from a queue.  The links present in a
// tile are processed ONE AT A TIME through a single 64-bit (depth key | triangle id) plane in shared memory:
//   raster : the link's triangles are taken 128 at a time, one per thread: setup -> a 64-byte record (edge
//            functions at the first candidate sample, per-pixel steps, clipped bbox).  A block scan of the
//            candidate-sample counts turns the batch into ONE flat sample space cut into 128 equal chunks, so
//            every thread tests the same number of samples whatever the triangle sizes are.  Covered samples
//            go to a per-warp queue (warp-ballot aggregated) that is drained with all 32 lanes busy: z/w from
//            the unsnapped clip positions, atomicMin into the plane.
//   pairs  : the plane is reduced to 35 row bitmasks; silhouette pixel pairs (covered next to empty) fall out
//            of XORs of those masks and are appended to the tile's pair list; their blend weights are computed
//            with all lanes busy and kept in the list (triangle, edge, alpha) -- this list is all the backward
//            needs, so the plane can be reused by the next link and nothing is ever re-rasterised.
//   gather : alpha is scattered to two small planes and every pixel adds colour + its (up to four) pair
//            contributions in the reference's order into the per-view sum.
// After the last link: S = min(sum, 1), mask write, (S - ref)^2, g = dL/dsum; then the backward walks the pair
// list link by link: analytic gradient of the active edge's two vertices, contracted with [x y z 1] on the
// fly, warp-shuffle reduced, added to d loss / d mvp[item, link] with fp64 atomics.
#ifndef EHB_RTHREADS
#define EHB_RTHREADS 256
#endif
#define EHB_RWARPS (EHB_RTHREADS / 32)
#define EHB_BATCH EHB_RTHREADS
#ifndef EHB_PAIRCAP
#define EHB_PAIRCAP 768      // pair-list entries kept in shared memory; the rest spills to a per-CTA global area
#endif

struct __align__(16) EhbRec {
    long long E[3];          // edge function minus its threshold at the first candidate sample (covered: >= 0)
    int ex[3], ey[3];        // edge vectors (1/16 px): one pixel right adds -16*ey, one pixel up adds +16*ex
    uint32_t geom;           // w | h << 8 | lx << 16 | ly << 24 : clipped bbox and its origin inside the plane
    uint32_t id;             // triangle id stored in the depth key
};

struct EhbPairEnt {
    uint32_t packed;         // idx (11) | d << 11 | own << 12 | side << 13 | di << 14
    uint32_t tri;
    float alpha;
};

struct __align__(16) EhbSmem {
    unsigned long long plane[EHB_NP];
    union {
        struct {
            EhbRec rec[EHB_BATCH];
            float clip[EHB_BATCH][12];
            int off[EHB_BATCH + 1];
        } r;
        struct {
            float alpha[2][EHB_NP];
        } a;
    } ov;
    float sum[EHB_NP + 3];
    EhbPairEnt pairs[EHB_PAIRCAP];
    unsigned long long cov[EHB_RS + 1], hx[EHB_RS + 1], vy[EHB_RS + 1];
    float mvp[EHB_MAX_LINKS * 16];
    int links[EHB_MAX_LINKS];
    uint32_t lstart[EHB_MAX_LINKS];
    uint32_t lcnt[EHB_MAX_LINKS];
    int segStart[EHB_MAX_LINKS + 1];
    int rowCnt[EHB_RS + 1];
    int warpTot[EHB_RWARPS];
    int nP;
    int work;
};

__device__ __forceinline__ unsigned long long ehb_bits(int lo, int hi)   // bits lo..hi (inclusive), empty if lo > hi
{
    lo = max(lo, 0); hi = min(hi, 63);
    if (lo > hi) return 0ull;
    const unsigned long long up = hi >= 63 ? ~0ull : ((1ull << (hi + 1)) - 1ull);
    return up & ~((1ull << lo) - 1ull);
}

__device__ __forceinline__ void ehb_shade_entry(uint32_t ent, EhbSmem& sm, int rx0, int ry0, float xs, float xo,
                                                float ys, float yo)
{
    const int t = ent >> 12, ly = (ent >> 6) & 63, lx = ent & 63;
    const float* c = sm.ov.r.clip[t];
    const float4 a = *reinterpret_cast<const float4*>(c), b = *reinterpret_cast<const float4*>(c + 4),
                 d = *reinterpret_cast<const float4*>(c + 8);
    const float p0[4] = {a.x, a.y, a.z, a.w}, p1[4] = {b.x, b.y, b.z, b.w}, p2[4] = {d.x, d.y, d.z, d.w};
    const float fx = xs * (float)(rx0 + lx) + xo, fy = ys * (float)(ry0 + ly) + yo;
    const float zw = ehb_shade_zw(p0, p1, p2, fx, fy);
    const unsigned long long key = ((unsigned long long)ehb_order_key(zw) << 32) | sm.ov.r.rec[t].id;
    unsigned long long* dst = sm.plane + ly * EHB_RS + lx;
    if (key < *dst) atomicMin(dst, key);
}

__device__ __forceinline__ float ehb_fast_div(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ double ehb_fast_div(double a, double b) { return a / b; }

// Exact span of covered samples in one row of a triangle's clipped bbox.  Edge k along the row is
// C_k(dx) = R_k + ax_k * dx (threshold folded in: covered <=> C_k >= 0 for all k), monotone in dx, so the
// covered set is an interval [lo, hi].  Each bound is estimated with one float division and then fixed up with
// exact integer evaluations, so the result equals the brute-force test of every sample.
template <typename I, typename F>
__device__ __forceinline__ void ehb_row_span(const I R0, const I R1, const I R2, const I ax0, const I ax1, const I ax2,
                                             int w, int& lo, int& hi)
{
    lo = 0; hi = w - 1;
    const I R[3] = {R0, R1, R2}, ax[3] = {ax0, ax1, ax2};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (ax[k] > 0) {
            if (R[k] < 0) {   // smallest dx with R + ax*dx >= 0
                int q = (int)fmin((F)w, ceil(ehb_fast_div((F)(-R[k]), (F)ax[k])));
                q = max(q, 0);
                while (q > 0 && R[k] + ax[k] * (I)(q - 1) >= 0) q--;
                while (q < w && R[k] + ax[k] * (I)q < 0) q++;
                lo = max(lo, q);
            }
        } else if (ax[k] < 0) {
            if (R[k] < 0) hi = -1;
            else {            // largest dx with R + ax*dx >= 0
                int q = (int)fmin((F)(w - 1), floor(ehb_fast_div((F)R[k], (F)(-ax[k]))));
                while (q < w - 1 && R[k] + ax[k] * (I)(q + 1) >= 0) q++;
                while (q >= 0 && R[k] + ax[k] * (I)q < 0) q--;
                hi = min(hi, q);
            }
        } else if (R[k] < 0) hi = -1;
    }
}

// Rows of the batch's triangles -> covered spans -> shaded samples.  The batch's rows form one flat space
// (off[] = prefix of the clipped bbox heights); each warp takes 32 rows at a time (one per lane), computes their
// spans, and the warp then shades the covered samples of those 32 rows cooperatively, 32 at a time.
template <typename I, typename F>
__device__ __forceinline__ void ehb_rows_fill(EhbSmem& sm, int nRows, int warp, int lane, int rx0, int ry0, float xs,
                                              float xo, float ys, float yo)
{
    for (int g = warp * 32; g < nRows; g += EHB_RWARPS * 32) {
        const int r = g + lane;
        int len = 0;
        uint32_t pos = 0;   // t << 12 | ly << 6 | lx of the first covered sample of this lane's row
        if (r < nRows) {
            int lo = 0, hi = EHB_BATCH;   // last t with off[t] <= r
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (sm.ov.r.off[mid] <= r) lo = mid; else hi = mid;
            }
            const int t = lo, dy = r - sm.ov.r.off[t];
            const EhbRec& rc = sm.ov.r.rec[t];
            const uint32_t gm = rc.geom;
            const int w = gm & 255, lx0 = (gm >> 16) & 255, ly0 = gm >> 24;
            const I R0 = (I)rc.E[0] + (I)16 * (I)rc.ex[0] * (I)dy, R1 = (I)rc.E[1] + (I)16 * (I)rc.ex[1] * (I)dy,
                    R2 = (I)rc.E[2] + (I)16 * (I)rc.ex[2] * (I)dy;
            int a, b;
            ehb_row_span<I, F>(R0, R1, R2, (I)-16 * (I)rc.ey[0], (I)-16 * (I)rc.ey[1], (I)-16 * (I)rc.ey[2], w, a, b);
            len = max(0, b - a + 1);
            pos = ((uint32_t)t << 12) | ((uint32_t)(ly0 + dy) << 6) | (uint32_t)(lx0 + a);
        }
        int inc = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        const int total = __shfl_sync(0xffffffffu, inc, 31);
        const int exc = inc - len;
        for (int j0 = 0; j0 < total; j0 += 32) {
            const int j = j0 + lane;
            int o = 0;   // owner lane = number of lanes whose inclusive prefix is <= j
#pragma unroll
            for (int st = 16; st >= 1; st >>= 1) {
                const int v = __shfl_sync(0xffffffffu, inc, o + st - 1);
                if (v <= j) o += st;
            }
            o = min(o, 31);
            const uint32_t opos = __shfl_sync(0xffffffffu, pos, o);
            const int oexc = __shfl_sync(0xffffffffu, exc, o);
            if (j < total) ehb_shade_entry(opos + (uint32_t)(j - oexc), sm, rx0, ry0, xs, xo, ys, yo);
        }
    }
}

#ifndef EHB_MIN_BLOCKS
#define EHB_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(EHB_RTHREADS, EHB_MIN_BLOCKS) ehb_k_raster(const __grid_constant__ EhbRobot rb,
                                                                             const __grid_constant__ EhbParams p)
{
    extern __shared__ __align__(16) unsigned char ehb_smem_raw[];
    EhbSmem& sm = *reinterpret_cast<EhbSmem*>(ehb_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = p.H, W = p.W;
    const bool perLink = p.Lk != 1;
    const float xs = 2.f / (float)W, xo = 1.f / (float)W - 1.f;
    const float ys = 2.f / (float)H, yo = 1.f / (float)H - 1.f;
    const int hlo = p.hlo;
    const bool needAA = p.mode == EHB_MODE_FUSED || p.mode == EHB_MODE_AA_FWD;
    const bool needPairs = needAA || p.mode == EHB_MODE_AA_BWD;
    const bool doBwd = (p.mode == EHB_MODE_FUSED && p.do_bwd) || p.mode == EHB_MODE_AA_BWD;
    // out region: interior, plus one column/row on the high side when the backward follows in this pass
    const int oext = (p.mode == EHB_MODE_FUSED && p.do_bwd) ? 1 : 0;
    const int ow = EHB_T + oext;
    EhbPairEnt* spill = p.pairSpill + (size_t)blockIdx.x * p.spillCap;
    auto pair_at = [&](int i) -> EhbPairEnt& { return i < EHB_PAIRCAP ? sm.pairs[i] : spill[i - EHB_PAIRCAP]; };

    // the alpha planes stay all-zero between uses (every scatter is undone after its gather)
    if (needAA)
        ;  // zeroed per link below: they share storage with the raster records

    for (;;) {
        if (tid == 0) {
            const unsigned w = atomicAdd(&p.ctr->workCursor, 1u);
            const unsigned nh = p.ctr->nHeavy, nl = p.ctr->nLight;
            sm.work = w < nh ? (int)w : (w < nh + nl ? (int)((unsigned)(p.items * p.ntiles) - 1u - (w - nh)) : -1);
        }
        __syncthreads();
        if (sm.work < 0) break;
        const uint32_t wid = p.tileList[sm.work];
        const int item = wid / p.ntiles, tile = wid - item * p.ntiles;
        const int tx = tile % p.ntx, ty = tile / p.ntx;
        const int x0 = tx * EHB_T, y0 = ty * EHB_T;
        const int rx0 = x0 - hlo, ry0 = y0 - hlo;
        const int rx1 = x0 + EHB_T - 1 + p.hhi, ry1 = y0 + EHB_T - 1 + p.hhi;
        const size_t bin0 = (size_t)wid * p.Lk;
        const size_t ibase = (size_t)item * H * W;

        if (warp == 0) {
            const uint32_t c = lane < p.Lk ? p.cnt[bin0 + lane] : 0u;
            const unsigned b = __ballot_sync(0xffffffffu, c > 0);
            if (c > 0) {
                const int k = __popc(b & ((1u << lane) - 1u));
                sm.links[k] = lane;
                sm.lstart[k] = p.start[bin0 + lane];
                sm.lcnt[k] = c;
            }
            if (lane == 0) { sm.nP = __popc(b); sm.segStart[0] = 0; }
        }
        for (int i = tid; i < p.L * 16; i += EHB_RTHREADS) sm.mvp[i] = __ldg(p.mvp + (size_t)item * p.L * 16 + i);
        if (needAA)
            for (int i = tid; i < EHB_NP; i += EHB_RTHREADS) sm.sum[i] = 0.f;
        if (p.mode == EHB_MODE_AA_BWD)   // g = dL/dmask comes from the caller
            for (int i = tid; i < (EHB_T + 1) * (EHB_T + 1); i += EHB_RTHREADS) {
                const int qy = i / (EHB_T + 1), qx = i - qy * (EHB_T + 1);
                const int px = x0 + qx, py = y0 + qy;
                if (px < W && py < H)
                    sm.sum[(py - ry0) * EHB_RS + (px - rx0)] = __ldg(p.dy + ibase + (size_t)(H - 1 - py) * W + px);
            }
        __syncthreads();
        const int nP = sm.nP;

        for (int k = 0; k < nP; k++) {
            const int l = sm.links[k];   // UNION: the single bin holds triangles of every link
            // ================================ raster of link k =================================================
            for (int i = tid; i < EHB_NP; i += EHB_RTHREADS) sm.plane[i] = EHB_EMPTY;
            const uint32_t lo = sm.lstart[k], n = sm.lcnt[k];
            for (uint32_t b0 = 0; b0 < n; b0 += EHB_BATCH) {
                __syncthreads();   // plane cleared / previous batch's records (or alpha planes) no longer in use
                int ns = 0, wide = 0;
                if (b0 + tid < n) {
                    const uint32_t e = __ldg(p.pairs + lo + b0 + tid);
                    const int el = e >> EHB_LINK_SHIFT;
                    const int f = e & EHB_FACE_MASK;
                    EhbTri s;
                    if (ehb_tri_setup<true>(rb.link[el], sm.mvp + 16 * el, f, H, W, s) == 0) {
                        const int xlo = max(s.pxlo, rx0), xhi = min(s.pxhi, rx1);
                        const int ylo = max(s.pylo, ry0), yhi = min(s.pyhi, ry1);
                        if (xlo <= xhi && ylo <= yhi) {
                            EhbRec& rc = sm.ov.r.rec[tid];
                            const int bx = 8 * W - 8, by = 8 * H - 8;
                            const int sx = 16 * xlo - bx, sy = 16 * ylo - by;
                            const int ex0 = s.x1 - s.x0, ey0 = s.y1 - s.y0, ex1 = s.x2 - s.x1, ey1 = s.y2 - s.y1,
                                      ex2 = s.x0 - s.x2, ey2 = s.y0 - s.y2;
                            rc.E[0] = (long long)ex0 * (sy - s.y0) - (long long)ey0 * (sx - s.x0) - (ehb_edge_inclusive(ex0, ey0, p.rule) ? 0 : 1);
                            rc.E[1] = (long long)ex1 * (sy - s.y1) - (long long)ey1 * (sx - s.x1) - (ehb_edge_inclusive(ex1, ey1, p.rule) ? 0 : 1);
                            rc.E[2] = (long long)ex2 * (sy - s.y2) - (long long)ey2 * (sx - s.x2) - (ehb_edge_inclusive(ex2, ey2, p.rule) ? 0 : 1);
                            rc.ex[0] = ex0; rc.ex[1] = ex1; rc.ex[2] = ex2;
                            rc.ey[0] = ey0; rc.ey[1] = ey1; rc.ey[2] = ey2;
                            rc.geom = (uint32_t)(xhi - xlo + 1) | ((uint32_t)(yhi - ylo + 1) << 8) |
                                      ((uint32_t)(xlo - rx0) << 16) | ((uint32_t)(ylo - ry0) << 24);
                            rc.id = perLink ? (uint32_t)f : (uint32_t)(rb.foff[el] + f);
                            float* cc = sm.ov.r.clip[tid];
                            *reinterpret_cast<float4*>(cc) = make_float4(s.c0[0], s.c0[1], s.c0[2], s.c0[3]);
                            *reinterpret_cast<float4*>(cc + 4) = make_float4(s.c1[0], s.c1[1], s.c1[2], s.c1[3]);
                            *reinterpret_cast<float4*>(cc + 8) = make_float4(s.c2[0], s.c2[1], s.c2[2], s.c2[3]);
                            ns = yhi - ylo + 1;   // rows of the clipped bbox
                            // 32-bit edge arithmetic is exact while every edge vector stays below 2^15 sub-pixel units
                            const int ext = max(max(abs(ex0), abs(ex1)), max(max(abs(ex2), abs(ey0)), max(abs(ey1), abs(ey2))));
                            wide = ext >= 32768;
                        }
                    }
                }
                int inc = ns;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                if (lane == 31) sm.warpTot[warp] = inc;
                const int anyWide = __syncthreads_or(wide);
                int wbase = 0, total = 0;
#pragma unroll
                for (int i = 0; i < EHB_RWARPS; i++) {
                    const int v = sm.warpTot[i];
                    if (i < warp) wbase += v;
                    total += v;
                }
                sm.ov.r.off[tid] = wbase + inc - ns;
                if (tid == 0) sm.ov.r.off[EHB_BATCH] = total;
                __syncthreads();
                if (anyWide) ehb_rows_fill<long long, double>(sm, total, warp, lane, rx0, ry0, xs, xo, ys, yo);
                else ehb_rows_fill<int, float>(sm, total, warp, lane, rx0, ry0, xs, xo, ys, yo);
            }
            __syncthreads();

            if (!needPairs) break;   // UNION: one pass, the plane is the result
            const EhbLink& lk = rb.link[l];
            const float* m = sm.mvp + 16 * l;
            // ================================ row bitmasks and silhouette pairs =================================
            for (int r = warp; r < EHB_RS; r += EHB_RWARPS) {
                const unsigned b0 = __ballot_sync(0xffffffffu, sm.plane[r * EHB_RS + lane] != EHB_EMPTY);
                const unsigned b1 = __ballot_sync(0xffffffffu, lane < EHB_RS - 32 && sm.plane[r * EHB_RS + 32 + lane] != EHB_EMPTY);
                if (lane == 0) sm.cov[r] = (unsigned long long)b0 | ((unsigned long long)b1 << 32);
            }
            if (tid == 0) sm.cov[EHB_RS] = 0ull;
            if (needAA)   // the alpha planes reuse the raster records' storage
                for (int i = tid; i < 2 * EHB_NP; i += EHB_RTHREADS) (&sm.ov.a.alpha[0][0])[i] = 0.f;
            __syncthreads();
            unsigned long long hxm = 0ull, vym = 0ull, ownm = 0ull;
            int ownRow = 0;
            if (tid < EHB_RS) {
                const int r = tid, py = ry0 + r;
                const unsigned long long cm = sm.cov[r], cu = sm.cov[r + 1];
                // columns whose pixel is inside the image, and for which the right neighbour is too
                const unsigned long long inX = ehb_bits(-rx0, W - 1 - rx0), inX1 = ehb_bits(-rx0, W - 2 - rx0);
                // pairs wanted: forward = those touching a pixel of the out region; otherwise only owned ones
                unsigned long long wantH, wantV;
                if (needAA) {
                    wantH = (r >= hlo && r <= hlo + ow - 1) ? ehb_bits(hlo - 1, hlo + ow - 1) : 0ull;
                    wantV = (r >= hlo - 1 && r <= hlo + ow - 1) ? ehb_bits(hlo, hlo + ow - 1) : 0ull;
                } else {
                    wantH = wantV = (r >= hlo && r <= hlo + EHB_T - 1) ? ehb_bits(hlo, hlo + EHB_T - 1) : 0ull;
                }
                const bool rowIn = py >= 0 && py < H;
                if (rowIn) hxm = (cm ^ (cm >> 1)) & inX1 & wantH & ehb_bits(0, EHB_RS - 2);
                if (rowIn && py < H - 1 && r < EHB_RS - 1) vym = (cm ^ cu) & inX & wantV;
                ownRow = r >= hlo && r <= hlo + EHB_T - 1;
                ownm = ownRow ? ehb_bits(hlo, hlo + EHB_T - 1) : 0ull;
                sm.hx[r] = hxm; sm.vy[r] = vym;
                sm.rowCnt[r] = __popcll(hxm) + __popcll(vym);
            }
            __syncthreads();
            const int seg0 = sm.segStart[k];
            if (tid < EHB_RS) {
                int o = seg0;
                for (int r = 0; r < tid; r++) o += sm.rowCnt[r];
                if (tid == EHB_RS - 1) sm.segStart[k + 1] = o + sm.rowCnt[tid];
                const uint32_t rowBase = (uint32_t)tid * EHB_RS;
                while (hxm) {
                    const int b = __ffsll((long long)hxm) - 1;
                    hxm &= hxm - 1;
                    pair_at(o++).packed = (rowBase + b) | (((ownm >> b) & 1ull) ? (1u << 12) : 0u);
                }
                while (vym) {
                    const int b = __ffsll((long long)vym) - 1;
                    vym &= vym - 1;
                    pair_at(o++).packed = (rowBase + b) | (1u << 11) | (((ownm >> b) & 1ull) ? (1u << 12) : 0u);
                }
            }
            __syncthreads();
            const int seg1 = sm.segStart[k + 1];
            // ================================ blend weights, all lanes busy =====================================
            for (int j = seg0 + tid; j < seg1; j += EHB_RTHREADS) {
                EhbPairEnt& pe = pair_at(j);
                const uint32_t pk = pe.packed;
                const int idx = pk & 2047, d = (pk >> 11) & 1;
                const int ly = idx / EHB_RS, lx = idx - ly * EHB_RS;
                const unsigned long long ka = sm.plane[idx], kb = sm.plane[idx + (d ? EHB_RS : 1)];
                const int side = ka != EHB_EMPTY ? 0 : 1;
                const uint32_t t = (uint32_t)(side ? kb : ka);
                int di;
                const float al = ehb_aa_pair(lk, m, (int)t, side, rx0 + lx, ry0 + ly, d, H, W, &di);
                pe.packed = pk | ((uint32_t)side << 13) | ((uint32_t)di << 14);
                pe.tri = t;
                pe.alpha = al;
                if (needAA) sm.ov.a.alpha[d][idx] = al;
            }
            if (needAA) {
                __syncthreads();
                // ============================ gather: sum += colour + pair contributions ========================
                // pixel (qx, qy) of the out region; column 32 (the extra one) is handled by the last pass
                for (int q = tid; q < ow * EHB_T + (oext ? ow : 0); q += EHB_RTHREADS) {
                    int qx, qy;
                    if (q < ow * EHB_T) { qy = q >> 5; qx = q & 31; } else { qy = q - ow * EHB_T; qx = EHB_T; }
                    const int lx = hlo + qx, ly = hlo + qy;
                    const unsigned long long cm = sm.cov[ly];
                    const bool c = (cm >> lx) & 1ull;
                    const bool h0 = (sm.hx[ly] >> lx) & 1ull, v0 = (sm.vy[ly] >> lx) & 1ull;
                    const bool h1 = (sm.hx[ly] >> (lx - 1)) & 1ull, v1 = (sm.vy[ly - 1] >> lx) & 1ull;
                    if (!(c | h0 | v0 | h1 | v1)) continue;
                    const int idx = ly * EHB_RS + lx;
                    const float cf = c ? 1.f : 0.f;
                    float o = cf, a;
                    // colour, pair(p,p+x), pair(p,p+y), pair(p-x,p), pair(p-y,p); the receiver is p0 when alpha > 0
                    if (h0) { a = sm.ov.a.alpha[0][idx]; if (a > 0.f) o += a * ((c ? 0.f : 1.f) - cf); }
                    if (v0) { a = sm.ov.a.alpha[1][idx]; if (a > 0.f) o += a * ((c ? 0.f : 1.f) - cf); }
                    if (h1) { a = sm.ov.a.alpha[0][idx - 1]; if (!(a > 0.f) && a != 0.f) o += a * (cf - (c ? 0.f : 1.f)); }
                    if (v1) { a = sm.ov.a.alpha[1][idx - EHB_RS]; if (!(a > 0.f) && a != 0.f) o += a * (cf - (c ? 0.f : 1.f)); }
                    sm.sum[idx] = sm.sum[idx] + o;   // links are added in link order (rb_solver.py:68); absent links add 0
                }
            }
            __syncthreads();
        }

        if (p.mode == EHB_MODE_UNION) {
            for (int i = tid; i < EHB_T * EHB_T; i += EHB_RTHREADS) {
                const int px = x0 + (i & 31), py = y0 + (i >> 5);
                if (px >= W || py >= H) continue;
                const unsigned long long kk = sm.plane[(py - ry0) * EHB_RS + (px - rx0)];
                p.out_u8[ibase + (size_t)(H - 1 - py) * W + px] =
                    (kk != EHB_EMPTY) && ((uint32_t)(kk >> 32) > 0x80000000u);
            }
        } else if (needAA) {
            // S = min(sum, 1); loss; g = dL/dsum kept in sm.sum for the backward
            double lacc = 0.0;
            const bool haveRef = p.ref != nullptr || p.ref_u8 != nullptr;
            for (int q = tid; q < ow * EHB_T + (oext ? ow : 0); q += EHB_RTHREADS) {
                int qx, qy;
                if (q < ow * EHB_T) { qy = q >> 5; qx = q & 31; } else { qy = q - ow * EHB_T; qx = EHB_T; }
                const int px = x0 + qx, py = y0 + qy;
                if (px >= W || py >= H) continue;
                const int idx = (py - ry0) * EHB_RS + (px - rx0);
                const float s = sm.sum[idx];
                const float S = (p.clamp && s > 1.f) ? 1.f : s;
                const size_t o = ibase + (size_t)(H - 1 - py) * W + px;
                const bool interior = qx < EHB_T && qy < EHB_T;
                if (interior && p.masks) p.masks[o] = S;
                if (haveRef) {
                    const float rf = p.ref ? __ldg(p.ref + o) : (__ldg(p.ref_u8 + o) ? 1.f : 0.f);
                    const float diff = S - rf;
                    if (interior) lacc += (double)(diff * diff);
                    sm.sum[idx] = (!p.clamp || s <= 1.f) ? (2.f * diff) * p.invB : 0.f;
                }
            }
            if (haveRef && p.loss) {
                lacc = ehb_warp_sum(lacc);
                if (lane == 0 && lacc != 0.0) atomicAdd(&p.loss[item], lacc);
            }
            __syncthreads();
        }

        // ======================================= backward: walk the pair list ===================================
        if (doBwd) {
            for (int k = 0; k < nP; k++) {
                const int l = sm.links[k];
                const EhbLink& lk = rb.link[l];
                const float* m = sm.mvp + 16 * l;
                const int seg0 = sm.segStart[k], seg1 = sm.segStart[k + 1];
                if (seg0 + warp * 32 >= seg1) continue;   // nothing for this warp (warp-uniform)
                double acc[12];
#pragma unroll
                for (int i = 0; i < 12; i++) acc[i] = 0.0;
                bool had = false;
                for (int j = seg0 + tid; j < seg1; j += EHB_RTHREADS) {
                    const EhbPairEnt pe = pair_at(j);
                    const float al = pe.alpha;
                    if (!(pe.packed & (1u << 12)) || al == 0.f) continue;
                    const int idx = pe.packed & 2047, d = (pe.packed >> 11) & 1, side = (pe.packed >> 13) & 1,
                              di = (pe.packed >> 14) & 3;
                    const int idx1 = idx + (d ? EHB_RS : 1);
                    const int ly = idx / EHB_RS, lx = idx - ly * EHB_RS;
                    const float g = sm.sum[al > 0.f ? idx : idx1];
                    const float dd = g * (side ? 1.f : -1.f);   // g * (c1 - c0)
                    if (dd == 0.f) continue;
                    int vi1, vi2;
                    float g1[3], g2[3];
                    ehb_aa_pair_grad(lk, m, (int)pe.tri, side, di, al, dd, rx0 + lx, ry0 + ly, d, H, W, &vi1, &vi2, g1, g2);
                    const float4 va = __ldg(lk.verts + vi1), vb = __ldg(lk.verts + vi2);
                    const double ha[4] = {(double)va.x, (double)va.y, (double)va.z, 1.0};
                    const double hb[4] = {(double)vb.x, (double)vb.y, (double)vb.z, 1.0};
#pragma unroll
                    for (int rr = 0; rr < 3; rr++)
#pragma unroll
                        for (int c = 0; c < 4; c++) acc[4 * rr + c] += (double)g1[rr] * ha[c] + (double)g2[rr] * hb[c];
                    had = true;
                    if (p.gpos) {
                        atomicAdd(p.gpos + 4 * (size_t)vi1 + 0, g1[0]);
                        atomicAdd(p.gpos + 4 * (size_t)vi1 + 1, g1[1]);
                        atomicAdd(p.gpos + 4 * (size_t)vi1 + 3, g1[2]);
                        atomicAdd(p.gpos + 4 * (size_t)vi2 + 0, g2[0]);
                        atomicAdd(p.gpos + 4 * (size_t)vi2 + 1, g2[1]);
                        atomicAdd(p.gpos + 4 * (size_t)vi2 + 3, g2[2]);
                    }
                }
                if (__any_sync(0xffffffffu, had)) {
                    double* dst = p.gmvp + ((size_t)item * p.L + l) * 16;
#pragma unroll
                    for (int i = 0; i < 12; i++) {
                        const double v = ehb_warp_sum(acc[i]);
                        // rows x (0), y (1), w (3) of d loss / d mvp; the z row carries no gradient
                        if (lane == 0 && v != 0.0) atomicAdd(dst + (i < 8 ? i : i + 4), v);
                    }
                }
            }
        }
        __syncthreads();
    }
}
