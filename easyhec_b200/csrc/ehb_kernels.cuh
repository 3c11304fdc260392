// ehb_kernels.cuh -- the kernels of one rasterizer pass.
//
// Work decomposition (B200: 148 SMs, 126 MB L2, 227 KB smem/SM).  The reference antialiases every link separately
// before summing (rb_solver.py:62-68), so visibility is resolved per (item, link) -- an "item" is a camera view or
// one render of a batch.  Each (item, link) gets a 64-bit (depth key | triangle id) plane covering only the
// link's screen bounding box; for robot views all planes of a step are a few tens of MB and stay L2-resident.
//   k_table  : one warp per plane: conservative screen bbox of the link from the projected corners of its <= 32
//              object-space chunk AABBs (256 point transforms); the last CTA to finish bump-allocates the planes
//   k_front  : three independent jobs in one launch: (a) one thread per (item, vertex): transform and snap ONCE (a
//              vertex is shared by ~6 triangles), keep the clip-space and snapped positions (a few MB, L2-resident);
//              (b) planes := EMPTY (allocated part only);  (c) one warp per 32x32 tile: tiles that some link's bbox
//              touches go to the tile queue (several links first), the others to the empty-tile list
//   k_raster : NO binning, NO block barriers: each warp takes 32 triangles (one per lane): setup -> record in the
//              warp's shared memory; the rows of the 32 clipped bboxes form one flat space; 32 rows at a time (one per
//              lane) get their exactly covered span (float estimate + integer fix-up == testing every sample), and
//              the warp then shades the covered samples of those rows cooperatively, 32 at a time: z/w from the
//              unsnapped clip positions, atomicMin (RED.MIN.U64, served by L2) into the plane.  Triangles that are not
//              small are deferred: record parked in global memory, bbox cut into 64x32 units for k_raster_big.
//              Spare CTAs of the same launch stream the empty tiles (mask = 0, loss += ref^2, float4): the HBM-bound
//              part of the frame overlaps the latency-bound part instead of preceding it.
//   k_tiles  : persistent CTAs over the non-empty tiles: per link whose bbox touches the tile: load the 35x35 window
//              of its plane, 35 row bitmasks -> silhouette pairs by XOR -> blend weights with all lanes busy -> pair
//              list (triangle, edge, alpha) -> gather into the per-view sum in the reference's order; then
//              S = min(sum, 1), mask write, (S - ref)^2, g = dL/dsum; the backward walks the pair list, contracts
//              the analytic vertex gradients with [x y z 1] on the fly, warp-shuffle reduce, fp64 atomicAdd into
//              d loss / d mvp[item, link].
// No intermediate image (rast, colour, antialias work queue, clip-space vertex buffer) is ever written.
#pragma once
#include "ehb_device.cuh"

#define EHB_T 32            // tile interior
#define EHB_RS 35           // window row stride = T + max halo (1 low, 2 high)
#define EHB_NP (EHB_RS * EHB_RS)

enum { EHB_MODE_FUSED = 0, EHB_MODE_AA_FWD = 1, EHB_MODE_AA_BWD = 2, EHB_MODE_UNION = 3 };

struct EhbPairEnt {
    uint32_t packed;         // idx (11) | d << 11 | own << 12 | side << 13 | di << 14
    uint32_t tri;
    float alpha;
};

struct EhbPlane {            // depth plane of one (item, link): pixels [x0, x0+w) x [y0, y0+h), GL rows
    int x0, y0, w, h;
    long long off;           // first element in the plane pool
    long long pad;
};

struct EhbUnit { uint32_t rec; unsigned short dx0, dy0; };   // a 64 x 32 pixel window of a deferred triangle's bbox

struct EhbCounters {
    unsigned long long planeCursor;
    unsigned long long nNeedClip;
    unsigned int nTiles;     // tiles touched by >= 2 link bboxes: listed from the front of tileList (served first) ...
    unsigned int nLight;     // ... the others from the back
    unsigned int workCursor;
    unsigned int nEmpty;     // tiles no link touches: streamed (mask = 0, loss += ref^2) by spare CTAs of the raster launch
    unsigned int vertexDone; // CTAs of k_table that have finished: the last one allocates the planes
    unsigned int nBigRec;    // deferred (not small) triangles: records parked in global memory ...
    unsigned int nUnits;     // ... and cut into bounded units that k_raster_big spreads over the whole chip
    unsigned int flags;      // 1: plane pool too small (results invalid, grow and rerun), 2: triangles need clipping
    unsigned int pad;
    unsigned long long dbg[16];   // EHB_TIMING builds: cycles per phase of k_tiles (thread 0 of every CTA)
};

struct EhbParams {
    int H, W, ntx, nty, ntiles;
    int items, L, Lp, Ftot, Vtot;   // Lp = planes per item: L (per-link visibility) or 1 (packed robot)
    int hlo, hhi;
    int mode, rule, do_bwd, clamp;
    float invB;
    const float* mvp;        // [items, L, 16]
    float4* vclip;           // [items, Vtot]  clip-space position of every vertex (written by k_front)
    int2* vsnap;             // [items, Vtot]  snapped screen position (1/16 px), x = INT_MIN when not drawable
    EhbPlane* plane;         // [items, Lp]
    unsigned long long* pool;
    unsigned long long poolCap;
    uint32_t* tileList;      // [items * ntiles]
    uint32_t* emptyList;     // [items * ntiles]
    uint32_t* touch;         // [items * ntiles]  bit l: a triangle of link l reaches into this tile's window
    struct EhbRec* bigRec;   // [bigCap]
    EhbUnit* units;          // [unitCap]
    int bigCap, unitCap;
    EhbCounters* ctr;
    const float* ref;        // [items, H, W]  FUSED
    const uint8_t* ref_u8;   // same, as bytes (either ref or ref_u8)
    float* masks;            // [items, H, W]  FUSED / AA_FWD
    double* loss;            // [items]
    double* gmvp;            // [items, L, 16]
    float* gpos;             // [V, 4] (AA_BWD, single link) or NULL
    const float* dy;         // [items, H, W]  AA_BWD
    uint8_t* out_u8;         // [items, H, W]  UNION
    EhbPairEnt* pairSpill;   // [gridDim.x of k_tiles][spillCap] overflow of the shared-memory pair lists
    int spillCap;
    unsigned long long* dbgbuf;   // EHB_TIMING builds: per k_tiles CTA {start ns, end ns, tiles, longest tile cycles, its links}
};

#ifndef EHB_SMALL_AREA
#define EHB_SMALL_AREA 96               // triangles whose clipped bbox has more candidate samples are deferred
#endif
#define EHB_UNIT_W 64
#define EHB_UNIT_H 32

// Programmatic dependent launch (sm_90+): the kernels of a pass are chained with the programmatic-stream-serialization
// launch attribute.  Each kernel first lets its successor be scheduled (its CTAs take free SM slots during this kernel's
// tail and run their prologue), then waits until its predecessor has completed and flushed.
__device__ __forceinline__ void ehb_pdl_enter()
{
#ifdef EHB_PDL
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

__device__ __forceinline__ int ehb_find_link(const int* off, int L, int g)
{
    int l = 0;
    while (l + 1 < L && g >= off[l + 1]) l++;
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

__device__ __forceinline__ double ehb_warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------ k_table
// One warp per depth plane: conservative screen bounding box of the link from the projected corners of its (up to
// 32) object-space chunk AABBs -- 256 point transforms instead of a reduction over all vertices, and no dependency on
// the vertex pass.  The last CTA to finish (atomic ticket) bump-allocates the planes in the pool.
__global__ void __launch_bounds__(256) ehb_k_table(const __grid_constant__ EhbRobot rb,
                                                   const __grid_constant__ EhbParams p)
{
    ehb_pdl_enter();
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        p.ctr->nTiles = 0u; p.ctr->nLight = 0u; p.ctr->nEmpty = 0u; p.ctr->workCursor = 0u;
        p.ctr->nBigRec = 0u; p.ctr->nUnits = 0u;
    }
    // outputs that the later kernels accumulate into
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.items; i += gridDim.x * blockDim.x) {
        if (p.loss) p.loss[i] = 0.0;
    }
    if (p.gmvp)
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.items * p.L * 16; i += gridDim.x * blockDim.x) p.gmvp[i] = 0.0;
    if (wid < p.items * p.Lp) {
        const int item = wid / p.Lp, pli = wid - item * p.Lp;
        float umin = 3.0e38f, umax = -3.0e38f, vmin = 3.0e38f, vmax = -3.0e38f;
        bool bad = false, any = false;
        const int l0 = p.Lp == 1 ? 0 : pli, l1 = p.Lp == 1 ? p.L : pli + 1;
        for (int l = l0; l < l1; l++) {
            const EhbLink& lk = rb.link[l];
            if (lane < lk.nboxes) {
                float m[16];
                ehb_load_mvp(p.mvp + ((size_t)item * p.L + l) * 16, m);
                const float4 lo = __ldg(lk.boxes + 2 * lane), hi = __ldg(lk.boxes + 2 * lane + 1);
                any = true;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const float4 v = make_float4((k & 1) ? hi.x : lo.x, (k & 2) ? hi.y : lo.y, (k & 4) ? hi.z : lo.z, 1.f);
                    float c[4];
                    ehb_xform(v, m, c);
                    if (!(c[3] > 1e-6f)) { bad = true; continue; }   // corner at or behind the camera plane: no finite bound
                    const float u = (c[0] / c[3] + 1.f) * (0.5f * (float)p.W), w = (c[1] / c[3] + 1.f) * (0.5f * (float)p.H);
                    umin = fminf(umin, u); umax = fmaxf(umax, u); vmin = fminf(vmin, w); vmax = fmaxf(vmax, w);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            umin = fminf(umin, __shfl_xor_sync(0xffffffffu, umin, o)); umax = fmaxf(umax, __shfl_xor_sync(0xffffffffu, umax, o));
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o)); vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        }
        bad = __any_sync(0xffffffffu, bad);
        any = __any_sync(0xffffffffu, any);
        if (lane == 0) {
            EhbPlane pl;
            pl.x0 = pl.y0 = pl.w = pl.h = 0; pl.off = 0; pl.pad = 0;
            if (any) {
                int x0 = 0, y0 = 0, x1 = p.W - 1, y1 = p.H - 1;
                if (!bad) {   // pixel p is sampled at p + 0.5; two pixels of margin cover snapping and rounding
                    x0 = max(0, (int)fmaxf(fminf(floorf(umin) - 2.f, 1e6f), -1e6f)); x1 = min(p.W - 1, (int)fmaxf(fminf(ceilf(umax) + 2.f, 1e6f), -1e6f));
                    y0 = max(0, (int)fmaxf(fminf(floorf(vmin) - 2.f, 1e6f), -1e6f)); y1 = min(p.H - 1, (int)fmaxf(fminf(ceilf(vmax) + 2.f, 1e6f), -1e6f));
                }
                if (x0 <= x1 && y0 <= y1) { pl.x0 = x0; pl.y0 = y0; pl.w = x1 - x0 + 1; pl.h = y1 - y0 + 1; }
            }
            p.plane[wid] = pl;
        }
    }
    // the last CTA to finish sees every bounding box: it bump-allocates the planes of this pass
    __shared__ unsigned s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&p.ctr->vertexDone, 1u) == gridDim.x - 1u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) { p.ctr->vertexDone = 0u; p.ctr->planeCursor = 0ull; }
    __syncthreads();
    for (int i = threadIdx.x; i < p.items * p.Lp; i += blockDim.x) {
        EhbPlane pl;
        {
            const int4 a = __ldcg(reinterpret_cast<const int4*>(&p.plane[i]));   // written by other CTAs of this launch: read at L2
            pl.x0 = a.x; pl.y0 = a.y; pl.w = a.z; pl.h = a.w; pl.off = 0; pl.pad = 0;
        }
        if (pl.w > 0) {
            const unsigned long long area = (unsigned long long)pl.w * (unsigned long long)pl.h;
            const unsigned long long off = atomicAdd(&p.ctr->planeCursor, area);
            if (off + area <= p.poolCap) pl.off = (long long)off;
            else { pl.w = pl.h = 0; atomicOr(&p.ctr->flags, 1u); }
            p.plane[i] = pl;
        }
    }
}

// ------------------------------------------------------------------------------------------------ empty tiles
// A tile no link touches: mask = 0 and loss += sum ref^2, streamed by one warp.
__device__ __forceinline__ void ehb_stream_empty_tile(const EhbParams& p, int item, int tx, int ty, int lane)
{
    const int x0 = tx * EHB_T, y0 = ty * EHB_T;
    const int H = p.H, W = p.W;
    const size_t ibase = (size_t)item * H * W;
    if (p.mode == EHB_MODE_FUSED || p.mode == EHB_MODE_AA_FWD) {
        double acc = 0.0;
        const bool vec = (W & 3) == 0 && ((((uintptr_t)p.masks) | ((uintptr_t)p.ref)) & 15) == 0 &&
                         (((uintptr_t)p.ref_u8) & 3) == 0;
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
    }
}

// ------------------------------------------------------------------------------------------------ k_front
// Three independent jobs in one launch (block ranges): (a) transform + snap every vertex once, (b) planes := EMPTY and
// link bitmaps := 0, (c) tile classification, one warp per tile.
__global__ void __launch_bounds__(256) ehb_k_front(const __grid_constant__ EhbRobot rb,
                                                   const __grid_constant__ EhbParams p, int vchunks, int clearBlocks)
{
    ehb_pdl_enter();
    const int vertexBlocks = vchunks * p.items;
    if ((int)blockIdx.x < vertexBlocks) {
        const int item = blockIdx.x / vchunks;
        const int g = (blockIdx.x - item * vchunks) * blockDim.x + threadIdx.x;
        if (g >= p.Vtot) return;
        const int lk = ehb_find_link(rb.voff, rb.L, g);
        float m[16], c[4];
        ehb_load_mvp(p.mvp + ((size_t)item * p.L + lk) * 16, m);
        ehb_xform(__ldg(rb.link[lk].verts + (g - rb.voff[lk])), m, c);
        int2 sn = make_int2(INT_MIN, 0);
        if (c[3] >= fabsf(c[2])) {   // only such vertices can belong to a drawable triangle
            const float r = 1.0f / c[3];
            sn = make_int2(ehb_rni_sat(c[0] * r * (float)(p.W * 8)), ehb_rni_sat(c[1] * r * (float)(p.H * 8)));
        }
        p.vclip[(size_t)item * p.Vtot + g] = make_float4(c[0], c[1], c[2], c[3]);
        p.vsnap[(size_t)item * p.Vtot + g] = sn;
        return;
    }
    const int cb = (int)blockIdx.x - vertexBlocks;
    if (cb < clearBlocks) {
        const unsigned long long total = min(p.ctr->planeCursor, p.poolCap);
        const unsigned long long n2 = total >> 1;
        ulonglong2* p2 = reinterpret_cast<ulonglong2*>(p.pool);
        for (unsigned long long i = (unsigned long long)cb * blockDim.x + threadIdx.x; i < n2;
             i += (unsigned long long)clearBlocks * blockDim.x)
            p2[i] = make_ulonglong2(EHB_EMPTY, EHB_EMPTY);
        if ((total & 1ull) && cb == 0 && threadIdx.x == 0) p.pool[total - 1] = EHB_EMPTY;
        if (p.touch)
            for (int i = cb * blockDim.x + threadIdx.x; i < p.items * p.ntiles; i += clearBlocks * blockDim.x) p.touch[i] = 0u;
        return;
    }
    const int wid = ((cb - clearBlocks) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= p.items * p.ntiles) return;
    const int item = wid / p.ntiles, tile = wid - item * p.ntiles;
    const int tx = tile % p.ntx, ty = tile / p.ntx;
    bool hit = false;
    if (lane < p.Lp) {
        const EhbPlane pl = p.plane[(size_t)item * p.Lp + lane];
        const int rx0 = tx * EHB_T - p.hlo, ry0 = ty * EHB_T - p.hlo;
        const int rx1 = tx * EHB_T + EHB_T - 1 + p.hhi, ry1 = ty * EHB_T + EHB_T - 1 + p.hhi;
        hit = pl.w > 0 && pl.x0 <= rx1 && pl.x0 + pl.w - 1 >= rx0 && pl.y0 <= ry1 && pl.y0 + pl.h - 1 >= ry0;
    }
    const unsigned hits = __ballot_sync(0xffffffffu, hit);
    if (lane != 0) return;
    if (hits) {
        // tiles with several links take several times longer in k_tiles: queue them first (front), the rest from the back
        if (__popc(hits) >= 2) p.tileList[atomicAdd(&p.ctr->nTiles, 1u)] = (uint32_t)wid;
        else p.tileList[(unsigned)(p.items * p.ntiles) - 1u - atomicAdd(&p.ctr->nLight, 1u)] = (uint32_t)wid;
    } else if (p.mode != EHB_MODE_AA_BWD) {
        p.emptyList[atomicAdd(&p.ctr->nEmpty, 1u)] = (uint32_t)wid;
    }
}

// ------------------------------------------------------------------------------------------------ k_raster
#ifndef EHB_RWARPS
#define EHB_RWARPS 8         // warps per raster CTA; every warp works alone on batches of 32 triangles
#endif

struct __align__(16) EhbRec {
    long long E[3];          // edge function minus its threshold at the first candidate sample (covered: >= 0)
    int ex[3], ey[3];        // edge vectors (1/16 px): one pixel right adds -16*ey, one pixel up adds +16*ex
    int x0, y0, w, h;        // clipped bbox (pixels, GL rows)
    long long base;          // pool index of pixel (0,0) of this triangle's plane: idx = base + py * pw + px
    int pw;                  // plane width
    uint32_t id;             // triangle id stored in the depth key
    float clip[12];          // unsnapped clip positions of the three vertices
};

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
                // the estimate is within 1e-3 of the true quotient (|q| <= w <= 8192), so one exact step fixes it
                if (q > 0 && R[k] + ax[k] * (I)(q - 1) >= 0) q--;
                else if (q < w && R[k] + ax[k] * (I)q < 0) q++;
                lo = max(lo, q);
            }
        } else if (ax[k] < 0) {
            if (R[k] < 0) hi = -1;
            else {            // largest dx with R + ax*dx >= 0
                int q = (int)fmin((F)(w - 1), floor(ehb_fast_div((F)R[k], (F)(-ax[k]))));
                if (q < w - 1 && R[k] + ax[k] * (I)(q + 1) >= 0) q++;
                else if (q >= 0 && R[k] + ax[k] * (I)q < 0) q--;
                hi = min(hi, q);
            }
        } else if (R[k] < 0) hi = -1;
    }
}

// Record word layout (32 x 32-bit words): E[3] (2 words each) | ex[3] | ey[3] | x0 y0 w h | base (2) | pw | id | clip[12].
// k_raster keeps the 32 records of a warp transposed in shared memory (word k of record t at [k*32 + t]: conflict-free
// for "every lane its own record" and for arbitrary t); k_raster_big keeps one record as is.
template <int TS, int KS>
struct EhbRecView {
    const uint32_t* b;
    __device__ __forceinline__ uint32_t u(int t, int k) const { return b[t * TS + k * KS]; }
    __device__ __forceinline__ int i(int t, int k) const { return (int)u(t, k); }
    __device__ __forceinline__ float f(int t, int k) const { return __uint_as_float(u(t, k)); }
    __device__ __forceinline__ long long ll(int t, int k) const
    {
        return (long long)(((unsigned long long)u(t, k + 1) << 32) | (unsigned long long)u(t, k));
    }
};
typedef EhbRecView<1, 32> EhbRecSoA;
typedef EhbRecView<32, 1> EhbRecAoS;

__device__ __forceinline__ void ehb_rec_store_soa(uint32_t* b, int t, const EhbRec& rc)
{
#pragma unroll
    for (int k = 0; k < 3; k++) {
        b[(2 * k) * 32 + t] = (uint32_t)(unsigned long long)rc.E[k];
        b[(2 * k + 1) * 32 + t] = (uint32_t)((unsigned long long)rc.E[k] >> 32);
        b[(6 + k) * 32 + t] = (uint32_t)rc.ex[k];
        b[(9 + k) * 32 + t] = (uint32_t)rc.ey[k];
    }
    b[12 * 32 + t] = (uint32_t)rc.x0; b[13 * 32 + t] = (uint32_t)rc.y0; b[14 * 32 + t] = (uint32_t)rc.w; b[15 * 32 + t] = (uint32_t)rc.h;
    b[16 * 32 + t] = (uint32_t)(unsigned long long)rc.base; b[17 * 32 + t] = (uint32_t)((unsigned long long)rc.base >> 32);
    b[18 * 32 + t] = (uint32_t)rc.pw; b[19 * 32 + t] = rc.id;
#pragma unroll
    for (int k = 0; k < 12; k++) b[(20 + k) * 32 + t] = __float_as_uint(rc.clip[k]);
}

// Primitive assembly of one triangle from the pre-transformed vertices -> record.  Same tests, in the same order,
// as ehb_tri_setup (the oracle's eho_rasterize).  Returns the number of rows of its clipped bbox (0: nothing to draw).
__device__ __forceinline__ int ehb_make_record(const EhbRobot& rb, const EhbParams& p, int item, int g, EhbRec& rc,
                                               int& link_out)
{
    const int l = ehb_find_link(rb.foff, rb.L, g);
    link_out = l;
    const int f = g - rb.foff[l];
    const EhbLink& lk = rb.link[l];
    const int4 id = __ldg(lk.faces + f);
    if ((unsigned)id.x >= (unsigned)lk.V || (unsigned)id.y >= (unsigned)lk.V || (unsigned)id.z >= (unsigned)lk.V) return 0;
    const size_t vb = (size_t)item * p.Vtot + rb.voff[l];
    const float4 c0 = p.vclip[vb + id.x], c1 = p.vclip[vb + id.y], c2 = p.vclip[vb + id.z];
    if ((c0.w < c0.x && c1.w < c1.x && c2.w < c2.x) || (c0.w < -c0.x && c1.w < -c1.x && c2.w < -c2.x) ||
        (c0.w < c0.y && c1.w < c1.y && c2.w < c2.y) || (c0.w < -c0.y && c1.w < -c1.y && c2.w < -c2.y) ||
        (c0.w < c0.z && c1.w < c1.z && c2.w < c2.z) || (c0.w < -c0.z && c1.w < -c1.z && c2.w < -c2.z))
        return 0;
    const int2 s0 = p.vsnap[vb + id.x], s1 = p.vsnap[vb + id.y], s2 = p.vsnap[vb + id.z];
    const int G = 1 << 28;
    if (s0.x == INT_MIN || s1.x == INT_MIN || s2.x == INT_MIN ||   // a vertex outside the depth range: needs the clipper
        s0.x > G || s0.x < -G || s0.y > G || s0.y < -G || s1.x > G || s1.x < -G || s1.y > G || s1.y < -G ||
        s2.x > G || s2.x < -G || s2.y > G || s2.y < -G) {
        atomicAdd(&p.ctr->nNeedClip, 1ull);
        atomicOr(&p.ctr->flags, 2u);
        return 0;
    }
    int x0 = s0.x, y0 = s0.y, x1 = s1.x, y1 = s1.y, x2 = s2.x, y2 = s2.y;
    const long long area = (long long)(x1 - x0) * (y2 - y0) - (long long)(y1 - y0) * (x2 - x0);
    if (area == 0) return 0;
    if (area < 0) { int t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
    const int bx = 8 * p.W - 8, by = 8 * p.H - 8;
    const int pxlo = max((min(x0, min(x1, x2)) + bx + 15) >> 4, 0), pxhi = min((max(x0, max(x1, x2)) + bx) >> 4, p.W - 1);
    const int pylo = max((min(y0, min(y1, y2)) + by + 15) >> 4, 0), pyhi = min((max(y0, max(y1, y2)) + by) >> 4, p.H - 1);
    if (pxlo > pxhi || pylo > pyhi) return 0;
    const EhbPlane pl = p.plane[(size_t)item * p.Lp + (p.Lp == 1 ? 0 : l)];
    if (pl.w == 0) return 0;   // pool overflow: flagged, the pass is rerun
    const int sx = 16 * pxlo - bx, sy = 16 * pylo - by;
    const int ex0 = x1 - x0, ey0 = y1 - y0, ex1 = x2 - x1, ey1 = y2 - y1, ex2 = x0 - x2, ey2 = y0 - y2;
    rc.E[0] = (long long)ex0 * (sy - y0) - (long long)ey0 * (sx - x0) - (ehb_edge_inclusive(ex0, ey0, p.rule) ? 0 : 1);
    rc.E[1] = (long long)ex1 * (sy - y1) - (long long)ey1 * (sx - x1) - (ehb_edge_inclusive(ex1, ey1, p.rule) ? 0 : 1);
    rc.E[2] = (long long)ex2 * (sy - y2) - (long long)ey2 * (sx - x2) - (ehb_edge_inclusive(ex2, ey2, p.rule) ? 0 : 1);
    rc.ex[0] = ex0; rc.ex[1] = ex1; rc.ex[2] = ex2;
    rc.ey[0] = ey0; rc.ey[1] = ey1; rc.ey[2] = ey2;
    rc.x0 = pxlo; rc.y0 = pylo; rc.w = pxhi - pxlo + 1; rc.h = pyhi - pylo + 1;
    rc.base = pl.off - (long long)pl.y0 * pl.w - pl.x0;
    rc.pw = pl.w;
    rc.id = p.Lp == 1 ? (uint32_t)g : (uint32_t)f;
    rc.clip[0] = c0.x; rc.clip[1] = c0.y; rc.clip[2] = c0.z; rc.clip[3] = c0.w;
    rc.clip[4] = c1.x; rc.clip[5] = c1.y; rc.clip[6] = c1.z; rc.clip[7] = c1.w;
    rc.clip[8] = c2.x; rc.clip[9] = c2.y; rc.clip[10] = c2.z; rc.clip[11] = c2.w;
    return rc.h;
}

template <class RV>
__device__ __forceinline__ void ehb_shade_global(const RV rv, int t, int px, int py, unsigned long long* pool, float xs,
                                                 float xo, float ys, float yo)
{
    const float p0[4] = {rv.f(t, 20), rv.f(t, 21), rv.f(t, 22), rv.f(t, 23)};
    const float p1[4] = {rv.f(t, 24), rv.f(t, 25), rv.f(t, 26), rv.f(t, 27)};
    const float p2[4] = {rv.f(t, 28), rv.f(t, 29), rv.f(t, 30), rv.f(t, 31)};
    const float fx = xs * (float)px + xo, fy = ys * (float)py + yo;
    const float zw = ehb_shade_zw(p0, p1, p2, fx, fy);
    const unsigned long long key = ((unsigned long long)ehb_order_key(zw) << 32) | rv.u(t, 19);
    atomicMin(pool + (rv.ll(t, 16) + (long long)py * rv.i(t, 18) + px), key);
}

// One group of up to 32 rows (one per lane): spans, then the warp shades the covered samples 32 at a time.
// `t` = this lane's record index in recs (-1: no row), `dy` = its row inside that record's bbox.
template <typename I, typename F, class RV>
__device__ __forceinline__ void ehb_rows_group(const RV rv, int t, int dy, int lane, unsigned long long* pool,
                                               float xs, float xo, float ys, float yo, int cx0 = 0, int cx1 = 1 << 20)
{
    int len = 0;
    uint32_t pos = 0;   // t << 26 | py << 13 | px of the first covered sample of this lane's row
    if (t >= 0) {
        const I R0 = (I)rv.ll(t, 0) + (I)16 * (I)rv.i(t, 6) * (I)dy, R1 = (I)rv.ll(t, 2) + (I)16 * (I)rv.i(t, 7) * (I)dy,
                R2 = (I)rv.ll(t, 4) + (I)16 * (I)rv.i(t, 8) * (I)dy;
        int a, b;
        ehb_row_span<I, F>(R0, R1, R2, (I)-16 * (I)rv.i(t, 9), (I)-16 * (I)rv.i(t, 10), (I)-16 * (I)rv.i(t, 11), rv.i(t, 14), a, b);
        a = max(a, cx0); b = min(b, cx1);   // window of a deferred triangle's unit
        len = max(0, b - a + 1);
        pos = ((uint32_t)t << 26) | ((uint32_t)(rv.i(t, 13) + dy) << 13) | (uint32_t)(rv.i(t, 12) + a);
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
        if (j < total) {
            const uint32_t q = opos + (uint32_t)(j - oexc);
            ehb_shade_global(rv, (int)(q >> 26), (int)(q & 8191u), (int)((q >> 13) & 8191u), pool, xs, xo, ys, yo);
        }
    }
}

__global__ void __launch_bounds__(EHB_RWARPS * 32) ehb_k_raster(const __grid_constant__ EhbRobot rb,
                                                                const __grid_constant__ EhbParams p, int streamBlocks,
                                                                int chunks)
{
    ehb_pdl_enter();
    __shared__ uint32_t s_rec[EHB_RWARPS][32 * 32];   // 32 records per warp, transposed
    __shared__ int s_off[EHB_RWARPS][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if ((int)blockIdx.x < streamBlocks) {
        // spare CTAs of this launch finish the tiles no link touches (mask = 0, loss += ref^2): pure HBM streaming that
        // overlaps the latency-bound rasterization instead of sitting in front of it
        const int n = (int)p.ctr->nEmpty;
        for (int i = blockIdx.x * EHB_RWARPS + warp; i < n; i += streamBlocks * EHB_RWARPS) {
            const int wid = (int)p.emptyList[i];
            const int item = wid / p.ntiles, tile = wid - item * p.ntiles;
            ehb_stream_empty_tile(p, item, tile % p.ntx, tile / p.ntx, lane);
        }
        return;
    }
    const int rb_ = (int)blockIdx.x - streamBlocks;
    const int item = rb_ / chunks;
    const int g = ((rb_ - item * chunks) * EHB_RWARPS + warp) * 32 + lane;
    const float xs = 2.f / (float)p.W, xo = 1.f / (float)p.W - 1.f;
    const float ys = 2.f / (float)p.H, yo = 1.f / (float)p.H - 1.f;
    const EhbRecSoA recs{s_rec[warp]};
    int* off = s_off[warp];
    int rows = 0, wide = 0;
    if (g < p.Ftot) {
        EhbRec rc;
        int link;
        rows = ehb_make_record(rb, p, item, g, rc, link);
        if (rows > 0 && p.touch) {   // tell k_tiles which (tile, link) windows this triangle reaches into
            const uint32_t bit = 1u << link;
            const int txlo = max(0, (rc.x0 - p.hhi) >> 5), txhi = min(p.ntx - 1, (rc.x0 + rc.w - 1 + p.hlo) >> 5);
            const int tylo = max(0, (rc.y0 - p.hhi) >> 5), tyhi = min(p.nty - 1, (rc.y0 + rc.h - 1 + p.hlo) >> 5);
            for (int ty = tylo; ty <= tyhi; ty++)
                for (int tx = txlo; tx <= txhi; tx++) {
                    uint32_t* w = p.touch + (size_t)item * p.ntiles + ty * p.ntx + tx;
                    if (!(__ldcg(w) & bit)) atomicOr(w, bit);   // mostly already set: a cached load instead of an atomic
                }
        }
        if (rows > 0) {
            const int ext = max(max(abs(rc.ex[0]), abs(rc.ex[1])), max(max(abs(rc.ex[2]), abs(rc.ey[0])), max(abs(rc.ey[1]), abs(rc.ey[2]))));
            wide = ext >= 32768;   // 32-bit edge arithmetic is exact below 2^15 sub-pixel units per edge
            if (rc.w * rc.h > EHB_SMALL_AREA) {   // not small: park the record, cut the bbox into bounded units
                const int nux = (rc.w + EHB_UNIT_W - 1) / EHB_UNIT_W, nuy = (rc.h + EHB_UNIT_H - 1) / EHB_UNIT_H;
                const unsigned k = atomicAdd(&p.ctr->nBigRec, 1u);
                const unsigned u0 = atomicAdd(&p.ctr->nUnits, (unsigned)(nux * nuy));
                if ((int)k < p.bigCap && (int)(u0 + nux * nuy) <= p.unitCap) {
                    p.bigRec[k] = rc;
                    for (int uy = 0; uy < nuy; uy++)
                        for (int ux = 0; ux < nux; ux++)
                            p.units[u0 + uy * nux + ux] = EhbUnit{k, (unsigned short)(ux * EHB_UNIT_W), (unsigned short)(uy * EHB_UNIT_H)};
                    rows = 0;
                } else {
                    atomicOr(&p.ctr->flags, 4u);   // queues full: this one is drawn inline (slow but complete); its units are void
                    for (int i = 0; i < nux * nuy && (int)(u0 + i) < p.unitCap; i++) p.units[u0 + i] = EhbUnit{0xFFFFFFFFu, 0, 0};
                }
            }
            if (rows > 0) ehb_rec_store_soa(s_rec[warp], lane, rc);
        }
    }
    int inc = rows;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    off[lane] = inc - rows;
    if (lane == 31) off[32] = inc;
    const bool anyWide = __any_sync(0xffffffffu, wide && rows > 0);
    __syncwarp();
    const int nRows = off[32];
    for (int r0 = 0; r0 < nRows; r0 += 32) {
        const int r = r0 + lane;
        int t = -1, dy = 0;
        if (r < nRows) {
            int lo = 0, hi = 32;   // last t with off[t] <= r
#pragma unroll
            for (int st = 0; st < 5; st++) {
                const int mid = (lo + hi) >> 1;
                if (off[mid] <= r) lo = mid; else hi = mid;
            }
            t = lo; dy = r - off[t];
        }
        if (anyWide) ehb_rows_group<long long, double, EhbRecSoA>(recs, t, dy, lane, p.pool, xs, xo, ys, yo);
        else ehb_rows_group<int, float, EhbRecSoA>(recs, t, dy, lane, p.pool, xs, xo, ys, yo);
    }
}

// Deferred triangles: warps stride over the unit list; one unit = a 64 x 32 pixel window of one triangle's bbox.
__global__ void __launch_bounds__(256) ehb_k_raster_big(const __grid_constant__ EhbParams p)
{
    ehb_pdl_enter();
    __shared__ EhbRec s_rec[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = min((int)p.ctr->nUnits, p.unitCap);
    const float xs = 2.f / (float)p.W, xo = 1.f / (float)p.W - 1.f;
    const float ys = 2.f / (float)p.H, yo = 1.f / (float)p.H - 1.f;
    EhbRec* rc = &s_rec[warp];
    for (int u = blockIdx.x * 8 + warp; u < n; u += gridDim.x * 8) {
        const EhbUnit un = p.units[u];
        if (un.rec == 0xFFFFFFFFu) continue;
        __syncwarp();
        reinterpret_cast<uint32_t*>(rc)[lane] = reinterpret_cast<const uint32_t*>(p.bigRec + un.rec)[lane];   // 128 B
        __syncwarp();
        const int ext = max(max(abs(rc->ex[0]), abs(rc->ex[1])), max(max(abs(rc->ex[2]), abs(rc->ey[0])), max(abs(rc->ey[1]), abs(rc->ey[2]))));
        const int dy = un.dy0 + lane;
        const int t = dy < rc->h ? 0 : -1;
        const EhbRecAoS rv{reinterpret_cast<const uint32_t*>(rc)};
        if (ext >= 32768) ehb_rows_group<long long, double, EhbRecAoS>(rv, t, dy, lane, p.pool, xs, xo, ys, yo, un.dx0, un.dx0 + EHB_UNIT_W - 1);
        else ehb_rows_group<int, float, EhbRecAoS>(rv, t, dy, lane, p.pool, xs, xo, ys, yo, un.dx0, un.dx0 + EHB_UNIT_W - 1);
    }
}

// UNION (packed robot, no antialiasing): mask = (z/w of the nearest triangle > 0), straight from the item's plane.
__global__ void __launch_bounds__(256) ehb_k_union_out(const __grid_constant__ EhbParams p)
{
    ehb_pdl_enter();
    const int item = blockIdx.y;
    const EhbPlane pl = p.plane[item];
    const int H = p.H, W = p.W;
    const size_t ibase = (size_t)item * H * W;
    const int nq = (W + 3) >> 2;   // groups of 4 pixels per row
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nq * H; i += gridDim.x * blockDim.x) {
        const int row = i / nq, q = i - row * nq;   // image row (row 0 = top)
        const int py = H - 1 - row;
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int px = 4 * q + k;
            if (px < W && px >= pl.x0 && px < pl.x0 + pl.w && py >= pl.y0 && py < pl.y0 + pl.h) {
                const unsigned long long key = p.pool[pl.off + (long long)(py - pl.y0) * pl.w + (px - pl.x0)];
                if (key != EHB_EMPTY && (uint32_t)(key >> 32) > 0x80000000u) v |= 1u << (8 * k);
            }
        }
        uint8_t* dst = p.out_u8 + ibase + (size_t)row * W + 4 * q;
        if (4 * q + 3 < W && (((uintptr_t)dst) & 3) == 0) *reinterpret_cast<uint32_t*>(dst) = v;
        else
            for (int k = 0; k < 4 && 4 * q + k < W; k++) dst[k] = (uint8_t)((v >> (8 * k)) & 1u);
    }
}

// ------------------------------------------------------------------------------------------------ k_tiles
#ifndef EHB_TTHREADS
#define EHB_TTHREADS 128
#endif
#define EHB_TWARPS (EHB_TTHREADS / 32)
static_assert(EHB_TTHREADS == 128, "k_tiles maps 32 rows x 4 eight-pixel segments onto its 128 threads");
#ifndef EHB_PAIRCAP
#define EHB_PAIRCAP 768      // pair-list entries kept in shared memory; the rest spills to a per-CTA global area
#endif

struct __align__(16) EhbSmem {
    unsigned long long plane[EHB_NP];
    float alpha[2][EHB_NP];
    float sum[EHB_NP + 3];
    EhbPairEnt pairs[EHB_PAIRCAP];
    unsigned long long cov[EHB_RS + 1], hx[EHB_RS + 1], vy[EHB_RS + 1];
    float mvp[EHB_MAX_LINKS * 16];
    EhbPlane pl[EHB_MAX_LINKS];
    int links[EHB_MAX_LINKS];
    int segStart[EHB_MAX_LINKS + 1];
    int rowCnt[EHB_RS + 1];
    int nP;
    int work;
    // reference mask of the out region, fetched with cp.async at the start of the tile so that its HBM latency is
    // hidden behind the antialias work: f32 words, or (u8 reference) 4 pixels per word
    uint32_t refw[(EHB_T + 1) * (EHB_T + 4)];
};

__device__ __forceinline__ void ehb_cp_async4(void* smem_dst, const void* gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void ehb_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ unsigned long long ehb_bits(int lo, int hi)   // bits lo..hi (inclusive), empty if lo > hi
{
    lo = max(lo, 0); hi = min(hi, 63);
    if (lo > hi) return 0ull;
    const unsigned long long up = hi >= 63 ? ~0ull : ((1ull << (hi + 1)) - 1ull);
    return up & ~((1ull << lo) - 1ull);
}

#ifdef EHB_TIMING
#define EHB_TICK(k) do { if (tid == 0) { const long long now_ = clock64(); atomicAdd(&p.ctr->dbg[k], (unsigned long long)(now_ - t_last)); t_last = now_; } } while (0)
#else
#define EHB_TICK(k) do { } while (0)
#endif
#ifndef EHB_TMIN_BLOCKS
#define EHB_TMIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(EHB_TTHREADS, EHB_TMIN_BLOCKS) ehb_k_tiles(const __grid_constant__ EhbRobot rb,
                                                                             const __grid_constant__ EhbParams p)
{
    ehb_pdl_enter();
    extern __shared__ __align__(16) unsigned char ehb_smem_raw[];
    EhbSmem& sm = *reinterpret_cast<EhbSmem*>(ehb_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = p.H, W = p.W;
    const int hlo = p.hlo;
    const bool needAA = p.mode == EHB_MODE_FUSED || p.mode == EHB_MODE_AA_FWD;
    const bool doBwd = (p.mode == EHB_MODE_FUSED && p.do_bwd) || p.mode == EHB_MODE_AA_BWD;
    // out region: interior, plus one column/row on the high side when the backward follows in this pass
    const int oext = (p.mode == EHB_MODE_FUSED && p.do_bwd) ? 1 : 0;
    const int ow = EHB_T + oext;
    EhbPairEnt* spill = p.pairSpill + (size_t)blockIdx.x * p.spillCap;
    auto pair_at = [&](int i) -> EhbPairEnt& { return i < EHB_PAIRCAP ? sm.pairs[i] : spill[i - EHB_PAIRCAP]; };

    // sm.alpha is never cleared: the gather only reads positions whose pair bit is set in THIS link's hx / vy masks,
    // and every such position is written by the blend-weight step of this link
    const bool haveRef = p.ref != nullptr || p.ref_u8 != nullptr;
    // u8 reference rows can be fetched 4 pixels per cp.async when every row start is 4-byte aligned
    const bool ref8vec = p.ref_u8 != nullptr && (p.W & 3) == 0 && (((uintptr_t)p.ref_u8) & 3) == 0;
    const unsigned nHeavy = p.ctr->nTiles, nTiles = nHeavy + p.ctr->nLight;
    const unsigned listEnd = (unsigned)(p.items * p.ntiles) - 1u;
    // thread 0 runs the tile queue two entries ahead: the atomic ticket and the tile id of the NEXT tile are fetched
    // while the current one is processed, so no tile starts with two dependent L2 round trips
    auto list_at = [&](unsigned w) -> uint32_t { return p.tileList[w < nHeavy ? w : listEnd - (w - nHeavy)]; };
    unsigned nextWork = 0, nextNext = 0;
    uint32_t nextWid = 0;
    if (tid == 0) {
        nextWork = atomicAdd(&p.ctr->workCursor, 1u);
        if (nextWork < nTiles) { nextWid = list_at(nextWork); nextNext = atomicAdd(&p.ctr->workCursor, 1u); }
    }

#ifdef EHB_TIMING
    long long t_last = clock64();
    unsigned long long t_cta0, n_done = 0, worst = 0, worstLinks = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_cta0));
#endif
    for (;;) {
        EHB_TICK(9);
#ifdef EHB_TIMING
        const long long t_tile0 = clock64();
#endif
        if (tid == 0) {
            sm.work = nextWork < nTiles ? (int)nextWid : -1;
            if (nextWork < nTiles) {
                nextWork = nextNext;
                if (nextWork < nTiles) { nextWid = list_at(nextWork); nextNext = atomicAdd(&p.ctr->workCursor, 1u); }
            }
        }
        __syncthreads();
        if (sm.work < 0) break;
        const uint32_t wid = (uint32_t)sm.work;
        const int item = wid / p.ntiles, tile = wid - item * p.ntiles;
        const int tx = tile % p.ntx, ty = tile / p.ntx;
        const int x0 = tx * EHB_T, y0 = ty * EHB_T;
        const int rx0 = x0 - hlo, ry0 = y0 - hlo;
        const int rx1 = x0 + EHB_T - 1 + p.hhi, ry1 = y0 + EHB_T - 1 + p.hhi;
        const size_t ibase = (size_t)item * H * W;

        if (warp == 0) {   // links whose plane touches this tile's window
            bool hit = false;
            EhbPlane pl;
            if (lane < p.L) {
                pl = p.plane[(size_t)item * p.L + lane];
                hit = pl.w > 0 && ((p.touch[wid] >> lane) & 1u);
            }
            const unsigned b = __ballot_sync(0xffffffffu, hit);
            if (hit) {
                const int k = __popc(b & ((1u << lane) - 1u));
                sm.links[k] = lane;
                sm.pl[k] = pl;
            }
            if (lane == 0) { sm.nP = __popc(b); sm.segStart[0] = 0; }
        }
        for (int i = tid; i < p.L * 16; i += EHB_TTHREADS) sm.mvp[i] = __ldg(p.mvp + (size_t)item * p.L * 16 + i);
        if (needAA && haveRef) {   // start fetching the reference mask of the out region
            if (p.ref) {
                for (int qy = warp; qy < ow; qy += EHB_TWARPS) {
                    const int py = y0 + qy;
                    if (py >= H) continue;
                    const float* row = p.ref + ibase + (size_t)(H - 1 - py) * W + x0;
                    if (x0 + lane < W) ehb_cp_async4(&sm.refw[qy * (EHB_T + 4) + lane], row + lane);
                    if (oext && lane == 0 && x0 + EHB_T < W) ehb_cp_async4(&sm.refw[qy * (EHB_T + 4) + EHB_T], row + EHB_T);
                }
            } else if (ref8vec) {
                for (int q = tid; q < ow * 16; q += EHB_TTHREADS) {
                    const int qy = q >> 4, qw = q & 15;
                    const int px = x0 + 4 * qw, py = y0 + qy;
                    if (qw < 9 && px < W && py < H && (qw < 8 || oext)) ehb_cp_async4(&sm.refw[qy * (EHB_T + 4) + qw], p.ref_u8 + ibase + (size_t)(H - 1 - py) * W + px);
                }
            }
        }
        if (needAA)
            for (int i = tid; i < (EHB_NP + 3) / 4; i += EHB_TTHREADS) reinterpret_cast<float4*>(sm.sum)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.mode == EHB_MODE_AA_BWD)   // g = dL/dmask comes from the caller
            for (int i = tid; i < (EHB_T + 1) * (EHB_T + 1); i += EHB_TTHREADS) {
                const int qy = i / (EHB_T + 1), qx = i - qy * (EHB_T + 1);
                const int px = x0 + qx, py = y0 + qy;
                if (px < W && py < H)
                    sm.sum[(py - ry0) * EHB_RS + (px - rx0)] = __ldg(p.dy + ibase + (size_t)(H - 1 - py) * W + px);
            }
        __syncthreads();
        const int nP = sm.nP;
        EHB_TICK(0);
#ifdef EHB_TIMING
        if (tid == 0) { atomicAdd(&p.ctr->dbg[10], 1ull); atomicAdd(&p.ctr->dbg[11], (unsigned long long)nP); }
#endif

        for (int k = 0; k < nP; k++) {
            const int l = sm.links[k];
            const EhbLink& lk = rb.link[l];
            const float* m = sm.mvp + 16 * l;
            // ================ window of link k's plane -> shared memory, row coverage masks by ballot ===============
            {
                const EhbPlane pl = sm.pl[k];
                unsigned long long anyRow = 0ull;
                // all loads of this warp's rows are issued before the first ballot consumes one (9 rows x 2 words in flight)
                constexpr int NR = (EHB_RS + EHB_TWARPS - 1) / EHB_TWARPS;
                unsigned long long v0[NR], v1[NR];
#pragma unroll
                for (int j = 0; j < NR; j++) {
                    const int r = warp + j * EHB_TWARPS;
                    v0[j] = EHB_EMPTY; v1[j] = EHB_EMPTY;
                    if (r < EHB_RS) {
                        const int py = ry0 + r - pl.y0;
                        const bool rowOk = py >= 0 && py < pl.h && ry0 + r <= ry1;
                        const unsigned long long* row = p.pool + pl.off + (long long)py * pl.w - pl.x0 + rx0;
                        const int px = rx0 + lane - pl.x0;
                        if (rowOk && px >= 0 && px < pl.w && rx0 + lane <= rx1) v0[j] = row[lane];
                        if (lane < EHB_RS - 32) {
                            const int px1 = px + 32;
                            if (rowOk && px1 >= 0 && px1 < pl.w && rx0 + 32 + lane <= rx1) v1[j] = row[32 + lane];
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < NR; j++) {
                    const int r = warp + j * EHB_TWARPS;
                    if (r < EHB_RS) {
                        sm.plane[r * EHB_RS + lane] = v0[j];
                        if (lane < EHB_RS - 32) sm.plane[r * EHB_RS + 32 + lane] = v1[j];
                        const unsigned b0 = __ballot_sync(0xffffffffu, v0[j] != EHB_EMPTY);
                        const unsigned b1 = __ballot_sync(0xffffffffu, v1[j] != EHB_EMPTY);
                        const unsigned long long cm = (unsigned long long)b0 | ((unsigned long long)b1 << 32);
                        if (lane == 0) sm.cov[r] = cm;
                        anyRow |= cm;
                    }
                }
                if (tid == 0) sm.cov[EHB_RS] = 0ull;
                if (!__syncthreads_or(anyRow != 0ull)) {   // the link reaches into the window but covers no sample of it
                    if (tid == 0) sm.segStart[k + 1] = sm.segStart[k];
                    __syncthreads();
                    EHB_TICK(1);
                    continue;
                }
            }
            EHB_TICK(1);
            const int seg0 = sm.segStart[k];
            if (warp == 0) {
                // columns whose pixel is inside the image, and for which the right neighbour is too
                const unsigned long long inX = ehb_bits(-rx0, W - 1 - rx0), inX1 = ehb_bits(-rx0, W - 2 - rx0);
                unsigned long long hm[2], vm[2], om[2];
                int cnt[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int r = lane + 32 * h;
                    hm[h] = vm[h] = om[h] = 0ull;
                    if (r < EHB_RS) {
                        const int py = ry0 + r;
                        const unsigned long long cm = sm.cov[r], cu = sm.cov[r + 1];
                        // pairs wanted: forward = those touching a pixel of the out region; otherwise only owned ones
                        unsigned long long wantH, wantV;
                        if (needAA) {
                            wantH = (r >= hlo && r <= hlo + ow - 1) ? ehb_bits(hlo - 1, hlo + ow - 1) : 0ull;
                            wantV = (r >= hlo - 1 && r <= hlo + ow - 1) ? ehb_bits(hlo, hlo + ow - 1) : 0ull;
                        } else {
                            wantH = wantV = (r >= hlo && r <= hlo + EHB_T - 1) ? ehb_bits(hlo, hlo + EHB_T - 1) : 0ull;
                        }
                        const bool rowIn = py >= 0 && py < H;
                        if (rowIn) hm[h] = (cm ^ (cm >> 1)) & inX1 & wantH & ehb_bits(0, EHB_RS - 2);
                        if (rowIn && py < H - 1 && r < EHB_RS - 1) vm[h] = (cm ^ cu) & inX & wantV;
                        om[h] = (r >= hlo && r <= hlo + EHB_T - 1) ? ehb_bits(hlo, hlo + EHB_T - 1) : 0ull;
                        sm.hx[r] = hm[h]; sm.vy[r] = vm[h];
                    }
                    cnt[h] = __popcll(hm[h]) + __popcll(vm[h]);
                }
                // exclusive prefix over the 35 rows: rows 0..31 by shuffle scan, rows 32..34 after them
                int inc = cnt[0];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                const int tot0 = __shfl_sync(0xffffffffu, inc, 31);
                int inc1 = cnt[1];
#pragma unroll
                for (int o = 1; o < 4; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, inc1, o);
                    if (lane >= o) inc1 += v;
                }
                const int tot1 = __shfl_sync(0xffffffffu, inc1, 3);
                if (lane == 0) sm.segStart[k + 1] = seg0 + tot0 + tot1;
                int o0 = seg0 + inc - cnt[0], o1 = seg0 + tot0 + inc1 - cnt[1];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    int o = h ? o1 : o0;
                    unsigned long long hxm = hm[h], vym = vm[h];
                    const uint32_t rowBase = (uint32_t)(lane + 32 * h) * EHB_RS;
                    while (hxm) {
                        const int b = __ffsll((long long)hxm) - 1;
                        hxm &= hxm - 1;
                        pair_at(o++).packed = (rowBase + b) | (((om[h] >> b) & 1ull) ? (1u << 12) : 0u);
                    }
                    while (vym) {
                        const int b = __ffsll((long long)vym) - 1;
                        vym &= vym - 1;
                        pair_at(o++).packed = (rowBase + b) | (1u << 11) | (((om[h] >> b) & 1ull) ? (1u << 12) : 0u);
                    }
                }
            }
            __syncthreads();
            const int seg1 = sm.segStart[k + 1];
            EHB_TICK(2);
#ifdef EHB_TIMING
            if (tid == 0) atomicAdd(&p.ctr->dbg[12], (unsigned long long)(seg1 - seg0));
#endif
            // ================================ blend weights, all lanes busy =====================================
            for (int j = seg0 + tid; j < seg1; j += EHB_TTHREADS) {
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
                if (needAA) sm.alpha[d][idx] = al;
            }
            if (needAA) {
                __syncthreads();
                EHB_TICK(3);
                // ============================ gather: sum += colour + pair contributions ========================
                // each thread owns 8 consecutive pixels of one row (32 rows x 4 segments); a second short pass takes the
                // extra row / column that exist when the backward follows.  Masks are read once per thread.
#ifndef EHB_SKIP_GATHER
                for (int pass = 0; pass < (oext ? 2 : 1); pass++) {
                    int qy, qx0, n;
                    if (pass == 0) { qy = tid >> 2; qx0 = (tid & 3) * 8; n = 8; }
                    else if (tid < 4) { qy = EHB_T; qx0 = tid * 8; n = 8; }            // row 32
                    else if (tid < 4 + EHB_T + 1) { qy = tid - 4; qx0 = EHB_T; n = 1; }   // column 32 (incl. the corner)
                    else break;
                    if (qy >= EHB_T + oext) continue;
                    const int ly = hlo + qy, lx0 = hlo + qx0;
                    const unsigned long long cm = sm.cov[ly], hr = sm.hx[ly], vr = sm.vy[ly], vd = sm.vy[ly - 1];
                    // the segment's 8 flag bits of each mask, extracted once
                    const uint32_t nm = (1u << n) - 1u;
                    const uint32_t c8 = (uint32_t)(cm >> lx0) & nm, h08 = (uint32_t)(hr >> lx0) & nm, v08 = (uint32_t)(vr >> lx0) & nm;
                    const uint32_t h18 = (uint32_t)(hr >> (lx0 - 1)) & nm, v18 = (uint32_t)(vd >> lx0) & nm;
                    const uint32_t pairBits = h08 | v08 | h18 | v18;
                    uint32_t todo = c8 | pairBits;
                    float* srow = sm.sum + ly * EHB_RS + lx0;
                    while (todo) {
                        const int i = __ffs(todo) - 1;
                        todo &= todo - 1;
                        const bool c = (c8 >> i) & 1u;
                        const float cf = c ? 1.f : 0.f;
                        float o = cf;
                        if ((pairBits >> i) & 1u) {
                            const int idx = ly * EHB_RS + lx0 + i;
                            const float nb = c ? 0.f : 1.f;   // a pair's other pixel has the opposite coverage
                            float a;
                            // colour, pair(p,p+x), pair(p,p+y), pair(p-x,p), pair(p-y,p); the receiver is p0 when alpha > 0
                            if ((h08 >> i) & 1u) { a = sm.alpha[0][idx]; if (a > 0.f) o += a * (nb - cf); }
                            if ((v08 >> i) & 1u) { a = sm.alpha[1][idx]; if (a > 0.f) o += a * (nb - cf); }
                            if ((h18 >> i) & 1u) { a = sm.alpha[0][idx - 1]; if (!(a > 0.f) && a != 0.f) o += a * (cf - nb); }
                            if ((v18 >> i) & 1u) { a = sm.alpha[1][idx - EHB_RS]; if (!(a > 0.f) && a != 0.f) o += a * (cf - nb); }
                        }
                        srow[i] = srow[i] + o;   // links are added in link order (rb_solver.py:68); absent links add 0
                    }
                }
#endif
            }
            __syncthreads();
            EHB_TICK(4);
        }

        if (needAA) {
            // S = min(sum, 1); loss; g = dL/dsum kept in sm.sum for the backward
            double lacc = 0.0;
            if (haveRef) { ehb_cp_async_wait_all(); __syncthreads(); }
            // same thread -> pixel mapping as the gather: 8 consecutive pixels of one row per thread, then the extra
            // row / column; the 8 mask values of a segment leave as two float4 stores
            const bool vecOut = p.masks != nullptr && (W & 3) == 0 && (((uintptr_t)p.masks) & 15) == 0;
            for (int pass = 0; pass < (oext ? 2 : 1); pass++) {
                int qy, qx0, n;
                if (pass == 0) { qy = tid >> 2; qx0 = (tid & 3) * 8; n = 8; }
                else if (tid < 4) { qy = EHB_T; qx0 = tid * 8; n = 8; }
                else if (tid < 4 + EHB_T + 1) { qy = tid - 4; qx0 = EHB_T; n = 1; }
                else break;
                const int py = y0 + qy;
                if (qy >= EHB_T + oext || py >= H) continue;
                const size_t orow = ibase + (size_t)(H - 1 - py) * W;
                float Sv[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    Sv[i] = 0.f;
                    const int qx = qx0 + i, px = x0 + qx;
                    if (i >= n || px >= W) continue;
                    const int idx = (py - ry0) * EHB_RS + (px - rx0);
                    const float s = sm.sum[idx];
                    const float S = (p.clamp && s > 1.f) ? 1.f : s;
                    Sv[i] = S;
                    const bool interior = qx < EHB_T && qy < EHB_T;
                    if (haveRef) {
                        float rf;
                        if (p.ref) rf = __uint_as_float(sm.refw[qy * (EHB_T + 4) + qx]);
                        else if (ref8vec) rf = ((sm.refw[qy * (EHB_T + 4) + (qx >> 2)] >> (8 * (qx & 3))) & 255u) ? 1.f : 0.f;
                        else rf = __ldg(p.ref_u8 + orow + px) ? 1.f : 0.f;
                        const float diff = S - rf;
                        if (interior) lacc += (double)(diff * diff);
                        sm.sum[idx] = (!p.clamp || s <= 1.f) ? (2.f * diff) * p.invB : 0.f;
                    }
                }
                if (p.masks && qy < EHB_T && qx0 < EHB_T) {
                    if (vecOut && x0 + qx0 + 7 < W) {
                        float4* dst = reinterpret_cast<float4*>(p.masks + orow + x0 + qx0);
                        dst[0] = make_float4(Sv[0], Sv[1], Sv[2], Sv[3]);
                        dst[1] = make_float4(Sv[4], Sv[5], Sv[6], Sv[7]);
                    } else {
                        for (int i = 0; i < n; i++)
                            if (x0 + qx0 + i < W) p.masks[orow + x0 + qx0 + i] = Sv[i];
                    }
                }
            }
            if (haveRef && p.loss) {
                lacc = ehb_warp_sum(lacc);
                if (lane == 0 && lacc != 0.0) atomicAdd(&p.loss[item], lacc);
            }
            __syncthreads();
        }
        EHB_TICK(5);

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
                for (int j = seg0 + tid; j < seg1; j += EHB_TTHREADS) {
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
        EHB_TICK(6);
#ifdef EHB_TIMING
        if (tid == 0 && p.dbgbuf) {
            unsigned long long t_now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
            unsigned long long* d = p.dbgbuf + (size_t)blockIdx.x * 5;
            d[0] = t_cta0; d[1] = t_now; d[2] = n_done + 1; d[3] = worst; d[4] = worstLinks;
        }
        if (tid == 0) {
            const unsigned long long dt = (unsigned long long)(clock64() - t_tile0);
            n_done++;
            if (dt > worst) { worst = dt; worstLinks = (unsigned long long)nP | ((unsigned long long)sm.segStart[nP] << 8); }
            atomicMax(&p.ctr->dbg[13], dt);
            if (dt > 50000ull) atomicAdd(&p.ctr->dbg[14], 1ull);
            if (dt > 100000ull) atomicAdd(&p.ctr->dbg[15], 1ull);
        }
#endif
    }
}
