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
#define EHB_THREADS 256

enum { EHB_MODE_FUSED = 0, EHB_MODE_AA_FWD = 1, EHB_MODE_AA_BWD = 2, EHB_MODE_UNION = 3, EHB_MODE_UNION_VAR = 4 };

struct EhbCounters {
    unsigned long long pairCursor;
    unsigned long long nNeedClip;
    unsigned int nTiles;
    unsigned int workCursor;
    unsigned int flags;
    unsigned int pad;
};

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
            p.ctr->nTiles = 0u;
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
    if (lane == 0) p.tileList[atomicAdd(&p.ctr->nTiles, 1u)] = (uint32_t)wid;
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
// One CTA per non-empty tile (persistent CTAs pull tiles from a queue).  Per tile:
//   raster : triangles are taken 256 at a time, one per thread: setup -> a 64-byte record in shared memory
//            (edge functions at the first candidate sample, per-pixel steps, clipped bbox).  A block scan of
//            the candidate-sample counts turns the batch into ONE flat sample space that is cut into 256 equal
//            chunks, so every thread tests the same number of samples whatever the triangle sizes are.
//            Covered samples are not shaded in place: they go to a per-warp queue (warp-ballot aggregated)
//            that is drained with all 32 lanes busy: z/w from the unsnapped clip positions, 64-bit
//            (depth key | triangle id) atomicMin into the link's plane.
//   AA fwd : per resident link: silhouette pixel pairs (covered next to empty) are compacted into a queue, their
//            blend weights are computed with all lanes busy and scattered into two alpha planes, then every
//            pixel adds colour + its four pair contributions in the reference's order.
//   loss   : S = min(sum_links, 1), (S - ref)^2, g = dL/dsum kept in shared memory.
//   AA bwd : pairs owned by interior pixels are compacted again, weights recomputed, analytic gradient of
//            the active edge's two vertices contracted with [x y z 1] on the fly, warp-shuffle reduced and
//            added to d loss / d mvp[item, link] with fp64 atomics.
struct __align__(16) EhbRec {
    long long E[3];          // edge functions at the first candidate sample of the clipped bbox
    int ex[3], ey[3];        // edge vectors (1/16 px): one pixel right adds -16*ey, one pixel up adds +16*ex
    unsigned short w, h;     // clipped bbox, in pixels
    unsigned char lx, ly;    // its origin inside the plane
    unsigned char slot, thr; // plane slot; bit k set: edge k excludes samples exactly on it
    uint32_t id;             // triangle id stored in the depth key
};

#define EHB_PLANES_BYTES(pmax) ((((size_t)(pmax) * EHB_NP * 8) + 15) & ~(size_t)15)
#define EHB_SUM_BYTES ((((size_t)EHB_NP * 4) + 15) & ~(size_t)15)
#define EHB_BATCH 256
#define EHB_WQ 128           // per-warp queue of covered samples
#define EHB_PAIRQ (2 * EHB_NP)

struct EhbRasterSmem {
    float mvp[EHB_MAX_LINKS * 16];
    int links[EHB_MAX_LINKS];     // present links of this tile, ascending
    uint32_t lstart[EHB_MAX_LINKS];
    uint32_t lcnt[EHB_MAX_LINKS];
    int slotOf[EHB_MAX_LINKS];    // link -> plane slot in the current round
    int warpTot[EHB_THREADS / 32];
    int nP;
    int work;
    unsigned int nTiles;
    int qn;                       // pair queue length
};

union EhbOverlay {
    struct {
        EhbRec rec[EHB_BATCH];
        float clip[EHB_BATCH][12];
        int off[EHB_BATCH + 1];
        uint32_t wq[EHB_THREADS / 32][EHB_WQ];
    } r;
    struct {
        float alpha[2][EHB_NP];
        uint32_t pairq[EHB_PAIRQ];
    } a;
};

__device__ __forceinline__ void ehb_shade_entry(uint32_t ent, const EhbOverlay& ov, unsigned long long* planes,
                                                int rx0, int ry0, float xs, float xo, float ys, float yo)
{
    const int t = ent >> 12, lx = (ent >> 6) & 63, ly = ent & 63;
    const float* c = ov.r.clip[t];
    const float4 a = *reinterpret_cast<const float4*>(c), b = *reinterpret_cast<const float4*>(c + 4),
                 d = *reinterpret_cast<const float4*>(c + 8);
    const float p0[4] = {a.x, a.y, a.z, a.w}, p1[4] = {b.x, b.y, b.z, b.w}, p2[4] = {d.x, d.y, d.z, d.w};
    const float fx = xs * (float)(rx0 + lx) + xo, fy = ys * (float)(ry0 + ly) + yo;
    const float zw = ehb_shade_zw(p0, p1, p2, fx, fy);
    const EhbRec& r = ov.r.rec[t];
    const unsigned long long key = ((unsigned long long)ehb_order_key(zw) << 32) | r.id;
    unsigned long long* dst = planes + r.slot * EHB_NP + ly * EHB_RS + lx;
    if (key < *dst) atomicMin(dst, key);
}

// Sweep this thread's chunk [s, s_end) of the batch's flat sample space; `iters` is the block-uniform trip count.
template <typename I>
__device__ __forceinline__ void ehb_sweep(const EhbOverlay& ov, uint32_t* wq, int& wqn, int s, int s_end, int iters,
                                          int lane, unsigned long long* planes, int rx0, int ry0, float xs, float xo,
                                          float ys, float yo)
{
    int t = 0, dx = 0, dy = 0, w = 1, rem = 0, lx0 = 0, ly0 = 0;
    I C0 = 0, C1 = 0, C2 = 0, R0 = 0, R1 = 0, R2 = 0, ax0 = 0, ax1 = 0, ax2 = 0, ay0 = 0, ay1 = 0, ay2 = 0;
    I t0 = 0, t1 = 0, t2 = 0;
    auto load = [&](int loc) {
        const EhbRec& r = ov.r.rec[t];
        w = r.w;
        lx0 = r.lx; ly0 = r.ly;
        dy = loc / w; dx = loc - dy * w;
        rem = (int)r.w * (int)r.h - loc;
        ax0 = (I)-16 * (I)r.ey[0]; ax1 = (I)-16 * (I)r.ey[1]; ax2 = (I)-16 * (I)r.ey[2];
        ay0 = (I)16 * (I)r.ex[0]; ay1 = (I)16 * (I)r.ex[1]; ay2 = (I)16 * (I)r.ex[2];
        R0 = (I)r.E[0] + ay0 * (I)dy; R1 = (I)r.E[1] + ay1 * (I)dy; R2 = (I)r.E[2] + ay2 * (I)dy;
        C0 = R0 + ax0 * (I)dx; C1 = R1 + ax1 * (I)dx; C2 = R2 + ax2 * (I)dx;
        t0 = r.thr & 1; t1 = (r.thr >> 1) & 1; t2 = (r.thr >> 2) & 1;
    };
    if (s < s_end) {
        int lo = 0, hi = EHB_BATCH;   // last t with off[t] <= s
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (ov.r.off[mid] <= s) lo = mid; else hi = mid;
        }
        t = lo;
        load(s - ov.r.off[t]);
    }
    for (int it = 0; it < iters; it++) {
        const bool act = s < s_end;
        const bool cov = act && C0 >= t0 && C1 >= t1 && C2 >= t2;
        const unsigned bal = __ballot_sync(0xffffffffu, cov);
        if (bal) {
            if (wqn + 32 > EHB_WQ) {   // drain the warp's queue with all lanes busy
                __syncwarp();
                for (int j = lane; j < wqn; j += 32) ehb_shade_entry(wq[j], ov, planes, rx0, ry0, xs, xo, ys, yo);
                __syncwarp();
                wqn = 0;
            }
            if (cov) wq[wqn + __popc(bal & ((1u << lane) - 1u))] = ((uint32_t)t << 12) | ((uint32_t)(lx0 + dx) << 6) | (uint32_t)(ly0 + dy);
            wqn += __popc(bal);
        }
        if (act) {
            s++; rem--; dx++;
            C0 += ax0; C1 += ax1; C2 += ax2;
            if (dx == w) { dx = 0; dy++; R0 += ay0; R1 += ay1; R2 += ay2; C0 = R0; C1 = R1; C2 = R2; }
            if (rem == 0 && s < s_end) {
                do { t++; } while (ov.r.off[t + 1] == ov.r.off[t]);
                load(0);
            }
        }
    }
}

template <int PMAX>
__global__ void __launch_bounds__(EHB_THREADS) ehb_k_raster(const __grid_constant__ EhbRobot rb,
                                                            const __grid_constant__ EhbParams p)
{
    extern __shared__ __align__(16) unsigned char ehb_smem_raw[];
    unsigned long long* planes = reinterpret_cast<unsigned long long*>(ehb_smem_raw);
    float* sumpl = reinterpret_cast<float*>(ehb_smem_raw + EHB_PLANES_BYTES(PMAX));
    EhbOverlay& ov = *reinterpret_cast<EhbOverlay*>(ehb_smem_raw + EHB_PLANES_BYTES(PMAX) + EHB_SUM_BYTES);
    EhbRasterSmem& sm = *reinterpret_cast<EhbRasterSmem*>(reinterpret_cast<unsigned char*>(&ov) + ((sizeof(EhbOverlay) + 15) & ~15));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = p.H, W = p.W;
    const bool perLink = p.Lk != 1;
    const float xs = 2.f / (float)W, xo = 1.f / (float)W - 1.f;
    const float ys = 2.f / (float)H, yo = 1.f / (float)H - 1.f;

    for (;;) {
        if (tid == 0) {
            sm.work = (int)atomicAdd(&p.ctr->workCursor, 1u);
            sm.nTiles = p.ctr->nTiles;
        }
        __syncthreads();
        if ((unsigned)sm.work >= sm.nTiles) break;
        const uint32_t wid = p.tileList[sm.work];
        const int item = wid / p.ntiles, tile = wid - item * p.ntiles;
        const int tx = tile % p.ntx, ty = tile / p.ntx;
        const int x0 = tx * EHB_T, y0 = ty * EHB_T;
        const int rx0 = x0 - p.hlo, ry0 = y0 - p.hlo;
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
            if (lane == 0) sm.nP = __popc(b);
        }
        for (int i = tid; i < p.L * 16; i += EHB_THREADS) sm.mvp[i] = __ldg(p.mvp + (size_t)item * p.L * 16 + i);
        __syncthreads();
        const int nP = sm.nP;
        const int nR = (nP + PMAX - 1) / PMAX;
        const bool needAA = p.mode == EHB_MODE_FUSED || p.mode == EHB_MODE_AA_FWD;
        // out region: interior, plus one column/row on the high side when the backward follows in this pass
        const int oext = (p.mode == EHB_MODE_FUSED && p.do_bwd) ? 1 : 0;
        const int ow = EHB_T + oext;

        // ---- one round = up to PMAX present links rasterised into the planes -------------------------------
        auto raster_round = [&](int r) {
            const int k0 = r * PMAX, k1 = min(nP, k0 + PMAX);
            for (int i = tid; i < (k1 - k0) * EHB_NP; i += EHB_THREADS) planes[i] = EHB_EMPTY;
            if (tid < EHB_MAX_LINKS) sm.slotOf[tid] = 0;
            __syncthreads();
            if (perLink && tid < k1 - k0) sm.slotOf[sm.links[k0 + tid]] = tid;
            const uint32_t lo = sm.lstart[k0];
            const uint32_t n = sm.lstart[k1 - 1] + sm.lcnt[k1 - 1] - lo;
            int wqn = 0;
            uint32_t* wq = ov.r.wq[warp];
            for (uint32_t b0 = 0; b0 < n; b0 += EHB_BATCH) {
                __syncthreads();   // slotOf visible / previous batch's records no longer in use
                // -- setup: one triangle per thread -> record
                int ns = 0, wide = 0;
                if (b0 + tid < n) {
                    const uint32_t e = __ldg(p.pairs + lo + b0 + tid);
                    const int l = e >> EHB_LINK_SHIFT;
                    const int f = e & EHB_FACE_MASK;
                    EhbTri s;
                    if (ehb_tri_setup(rb.link[l], sm.mvp + 16 * l, f, H, W, s) == 0) {
                        const int xlo = max(s.pxlo, rx0), xhi = min(s.pxhi, rx1);
                        const int ylo = max(s.pylo, ry0), yhi = min(s.pyhi, ry1);
                        if (xlo <= xhi && ylo <= yhi) {
                            EhbRec& rc = ov.r.rec[tid];
                            const int bx = 8 * W - 8, by = 8 * H - 8;
                            const int sx = 16 * xlo - bx, sy = 16 * ylo - by;
                            const int ex0 = s.x1 - s.x0, ey0 = s.y1 - s.y0, ex1 = s.x2 - s.x1, ey1 = s.y2 - s.y1,
                                      ex2 = s.x0 - s.x2, ey2 = s.y0 - s.y2;
                            rc.E[0] = (long long)ex0 * (sy - s.y0) - (long long)ey0 * (sx - s.x0);
                            rc.E[1] = (long long)ex1 * (sy - s.y1) - (long long)ey1 * (sx - s.x1);
                            rc.E[2] = (long long)ex2 * (sy - s.y2) - (long long)ey2 * (sx - s.x2);
                            rc.ex[0] = ex0; rc.ex[1] = ex1; rc.ex[2] = ex2;
                            rc.ey[0] = ey0; rc.ey[1] = ey1; rc.ey[2] = ey2;
                            rc.w = (unsigned short)(xhi - xlo + 1); rc.h = (unsigned short)(yhi - ylo + 1);
                            rc.lx = (unsigned char)(xlo - rx0); rc.ly = (unsigned char)(ylo - ry0);
                            rc.slot = (unsigned char)sm.slotOf[l];
                            rc.thr = (unsigned char)((ehb_edge_inclusive(ex0, ey0, p.rule) ? 0 : 1) |
                                                     (ehb_edge_inclusive(ex1, ey1, p.rule) ? 0 : 2) |
                                                     (ehb_edge_inclusive(ex2, ey2, p.rule) ? 0 : 4));
                            rc.id = perLink ? (uint32_t)f : (uint32_t)(rb.foff[l] + f);
                            float* cc = ov.r.clip[tid];
#pragma unroll
                            for (int i = 0; i < 4; i++) { cc[i] = s.c0[i]; cc[4 + i] = s.c1[i]; cc[8 + i] = s.c2[i]; }
                            ns = (xhi - xlo + 1) * (yhi - ylo + 1);
                            // 32-bit edge arithmetic is exact while every edge vector stays below 2^15 sub-pixel units
                            const int ext = max(max(abs(ex0), abs(ex1)), max(max(abs(ex2), abs(ey0)), max(abs(ey1), abs(ey2))));
                            wide = ext >= 32768;
                        }
                    }
                }
                // -- block scan of the candidate-sample counts
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
                for (int i = 0; i < EHB_THREADS / 32; i++) {
                    const int v = sm.warpTot[i];
                    if (i < warp) wbase += v;
                    total += v;
                }
                ov.r.off[tid] = wbase + inc - ns;
                if (tid == 0) ov.r.off[EHB_BATCH] = total;
                __syncthreads();
                // -- balanced sweep of the flat sample space
                const int chunk = (total + EHB_THREADS - 1) / EHB_THREADS;
                const int s0 = min(tid * chunk, total), s1 = min(s0 + chunk, total);
                if (anyWide) ehb_sweep<long long>(ov, wq, wqn, s0, s1, chunk, lane, planes, rx0, ry0, xs, xo, ys, yo);
                else ehb_sweep<int>(ov, wq, wqn, s0, s1, chunk, lane, planes, rx0, ry0, xs, xo, ys, yo);
                __syncwarp();
                for (int j = lane; j < wqn; j += 32) ehb_shade_entry(wq[j], ov, planes, rx0, ry0, xs, xo, ys, yo);
                __syncwarp();
                wqn = 0;
            }
            __syncthreads();
        };

        // ---- silhouette pairs of one plane -> queue.  Pair (q, d): q and its right (d=0) / upper (d=1) neighbour,
        //      one covered and one empty.  fwd: every pair of the region (and alpha planes zeroed); bwd: pairs
        //      whose first pixel is an interior pixel.
        auto detect_pairs = [&](const unsigned long long* pl, bool fwd) {
            if (tid == 0) sm.qn = 0;
            __syncthreads();
            const int rw = fwd ? (rx1 - rx0) : EHB_T, rh = fwd ? (ry1 - ry0) : EHB_T;   // first pixels visited
            const int bx = fwd ? rx0 : x0, by = fwd ? ry0 : y0;
            const int total = rw * rh;
            for (int i0 = 0; i0 < total; i0 += EHB_THREADS) {
                const int i = i0 + tid;
                unsigned m = 0;
                int idx = 0;
                if (i < total) {
                    const int qy = i / rw, qx = i - qy * rw;
                    const int px = bx + qx, py = by + qy;
                    idx = (py - ry0) * EHB_RS + (px - rx0);
                    if (fwd) { ov.a.alpha[0][idx] = 0.f; ov.a.alpha[1][idx] = 0.f; }
                    if (px >= 0 && py >= 0 && px < W && py < H) {
                        const bool c0 = pl[idx] != EHB_EMPTY;
                        if (px < W - 1 && (pl[idx + 1] != EHB_EMPTY) != c0) m |= 1u;
                        if (py < H - 1 && (pl[idx + EHB_RS] != EHB_EMPTY) != c0) m |= 2u;
                    }
                }
                const unsigned b0 = __ballot_sync(0xffffffffu, m & 1u), b1 = __ballot_sync(0xffffffffu, m & 2u);
                const int n0 = __popc(b0), n1 = __popc(b1);
                int base = 0;
                if (lane == 0 && n0 + n1) base = atomicAdd(&sm.qn, n0 + n1);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (m & 1u) ov.a.pairq[base + __popc(b0 & ((1u << lane) - 1u))] = (uint32_t)idx << 1;
                if (m & 2u) ov.a.pairq[base + n0 + __popc(b1 & ((1u << lane) - 1u))] = ((uint32_t)idx << 1) | 1u;
            }
            __syncthreads();
        };

        // ---- antialias forward of the links resident in the planes: sumpl += per-link antialiased value ------
        auto aa_forward_round = [&](int r) {
            const int k0 = r * PMAX, k1 = min(nP, k0 + PMAX);
            for (int k = k0; k < k1; k++) {
                const int l = sm.links[k];
                const EhbLink& lk = rb.link[l];
                const float* m = sm.mvp + 16 * l;
                const unsigned long long* pl = planes + (k - k0) * EHB_NP;
                detect_pairs(pl, true);
                const int qn = sm.qn;
                for (int j = tid; j < qn; j += EHB_THREADS) {
                    const uint32_t e = ov.a.pairq[j];
                    const int d = e & 1, idx = e >> 1;
                    const int ly = idx / EHB_RS, lx = idx - ly * EHB_RS;
                    const unsigned long long ka = pl[idx], kb = pl[idx + (d ? EHB_RS : 1)];
                    const bool c0 = ka != EHB_EMPTY;
                    int di;
                    ov.a.alpha[d][idx] = ehb_aa_pair(lk, m, (int)(uint32_t)(c0 ? ka : kb), c0 ? 0 : 1, rx0 + lx, ry0 + ly,
                                                     d, H, W, &di);
                }
                __syncthreads();
                for (int i = tid; i < ow * ow; i += EHB_THREADS) {
                    const int qy = i / ow, qx = i - qy * ow;
                    const int px = x0 + qx, py = y0 + qy;
                    if (px >= W || py >= H) continue;
                    const int idx = (py - ry0) * EHB_RS + (px - rx0);
                    const float cf = pl[idx] != EHB_EMPTY ? 1.f : 0.f;
                    float o = cf, a;
                    // colour, pair(p,p+x), pair(p,p+y), pair(p-x,p), pair(p-y,p): the receiving pixel is p0 when alpha > 0
                    a = ov.a.alpha[0][idx];
                    if (a > 0.f) o += a * ((pl[idx + 1] != EHB_EMPTY ? 1.f : 0.f) - cf);
                    a = ov.a.alpha[1][idx];
                    if (a > 0.f) o += a * ((pl[idx + EHB_RS] != EHB_EMPTY ? 1.f : 0.f) - cf);
                    a = ov.a.alpha[0][idx - 1];
                    if (!(a > 0.f) && a != 0.f) o += a * (cf - (pl[idx - 1] != EHB_EMPTY ? 1.f : 0.f));
                    a = ov.a.alpha[1][idx - EHB_RS];
                    if (!(a > 0.f) && a != 0.f) o += a * (cf - (pl[idx - EHB_RS] != EHB_EMPTY ? 1.f : 0.f));
                    sumpl[idx] = sumpl[idx] + o;   // links are added in link order (rb_solver.py:68); absent links add 0
                }
                __syncthreads();
            }
        };

        // ---- antialias backward of the links resident in the planes ----------------------------------------
        auto backward_round = [&](int r) {
            const int k0 = r * PMAX, k1 = min(nP, k0 + PMAX);
            for (int k = k0; k < k1; k++) {
                const int l = sm.links[k];
                const EhbLink& lk = rb.link[l];
                const float* m = sm.mvp + 16 * l;
                const unsigned long long* pl = planes + (k - k0) * EHB_NP;
                detect_pairs(pl, false);
                const int qn = sm.qn;
                double acc[12];
#pragma unroll
                for (int i = 0; i < 12; i++) acc[i] = 0.0;
                bool had = false;
                for (int j = tid; j < qn; j += EHB_THREADS) {
                    const uint32_t e = ov.a.pairq[j];
                    const int d = e & 1, idx = e >> 1, idx1 = idx + (d ? EHB_RS : 1);
                    const int ly = idx / EHB_RS, lx = idx - ly * EHB_RS;
                    const int px = rx0 + lx, py = ry0 + ly;
                    const unsigned long long ka = pl[idx], kb = pl[idx1];
                    const bool c0 = ka != EHB_EMPTY;
                    const int side = c0 ? 0 : 1;
                    const int t = (int)(uint32_t)(c0 ? ka : kb);
                    int di;
                    const float al = ehb_aa_pair(lk, m, t, side, px, py, d, H, W, &di);
                    if (al == 0.f) continue;
                    const float g = sumpl[al > 0.f ? idx : idx1];
                    const float dd = g * (c0 ? -1.f : 1.f);   // g * (c1 - c0)
                    if (dd == 0.f) continue;
                    int vi1, vi2;
                    float g1[3], g2[3];
                    ehb_aa_pair_grad(lk, m, t, side, di, al, dd, px, py, d, H, W, &vi1, &vi2, g1, g2);
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
                __syncthreads();
            }
        };

        // ======================================= forward =====================================================
        if (needAA)
            for (int i = tid; i < EHB_NP; i += EHB_THREADS) sumpl[i] = 0.f;
        for (int r = 0; r < nR; r++) {
            raster_round(r);
            if (needAA) aa_forward_round(r);
        }

        if (p.mode == EHB_MODE_UNION) {
            for (int i = tid; i < EHB_T * EHB_T; i += EHB_THREADS) {
                const int px = x0 + (i & 31), py = y0 + (i >> 5);
                if (px >= W || py >= H) continue;
                const unsigned long long k = planes[(py - ry0) * EHB_RS + (px - rx0)];
                p.out_u8[ibase + (size_t)(H - 1 - py) * W + px] =
                    (k != EHB_EMPTY) && ((uint32_t)(k >> 32) > 0x80000000u);
            }
        } else if (needAA) {
            // S = min(sum, 1); loss; g = dL/dsum kept in sumpl for the backward
            double lacc = 0.0;
            const bool haveRef = p.ref != nullptr || p.ref_u8 != nullptr;
            for (int i = tid; i < ow * ow; i += EHB_THREADS) {
                const int qy = i / ow, qx = i - qy * ow;
                const int px = x0 + qx, py = y0 + qy;
                if (px >= W || py >= H) continue;
                const int idx = (py - ry0) * EHB_RS + (px - rx0);
                const float s = sumpl[idx];
                const float S = (p.clamp && s > 1.f) ? 1.f : s;
                const size_t o = ibase + (size_t)(H - 1 - py) * W + px;
                const bool interior = qx < EHB_T && qy < EHB_T;
                if (interior && p.masks) p.masks[o] = S;
                if (haveRef) {
                    const float rf = p.ref ? __ldg(p.ref + o) : (__ldg(p.ref_u8 + o) ? 1.f : 0.f);
                    const float diff = S - rf;
                    if (interior) lacc += (double)(diff * diff);
                    sumpl[idx] = (!p.clamp || s <= 1.f) ? (2.f * diff) * p.invB : 0.f;
                }
            }
            if (haveRef && p.loss) {
                lacc = ehb_warp_sum(lacc);
                if (lane == 0 && lacc != 0.0) atomicAdd(&p.loss[item], lacc);
            }
            __syncthreads();
        } else if (p.mode == EHB_MODE_AA_BWD) {
            for (int i = tid; i < (EHB_T + 1) * (EHB_T + 1); i += EHB_THREADS) {
                const int qx = i % (EHB_T + 1), qy = i / (EHB_T + 1);
                const int px = x0 + qx, py = y0 + qy;
                if (px >= W || py >= H) continue;
                sumpl[(py - ry0) * EHB_RS + (px - rx0)] = __ldg(p.dy + ibase + (size_t)(H - 1 - py) * W + px);
            }
            __syncthreads();
        }

        // ======================================= backward ====================================================
        if ((p.mode == EHB_MODE_FUSED && p.do_bwd) || p.mode == EHB_MODE_AA_BWD) {
            for (int r = 0; r < nR; r++) {
                if (nR > 1) raster_round(r);
                backward_round(r);
            }
        }
        __syncthreads();
    }
}
