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
#define EHB_BIGQ 512
#define EHB_BIG_SAMPLES 64  // a triangle with more candidate samples in the tile goes to the warp path

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
struct EhbTileCtx {
    int item, tx, ty;
    int rx0, ry0, rx1, ry1;  // pixel extent of the region held in the planes (clipped to the image later)
};

template <typename I>
__device__ __forceinline__ void ehb_cover_rows(const EhbTri& s, int xlo, int xhi, int ylo, int yhi,
                                               unsigned long long* pl, int rx0, int ry0, uint32_t id, int H, int W,
                                               int rule)
{
    const int bx = 8 * W - 8, by = 8 * H - 8;
    const int ex0 = s.x1 - s.x0, ey0 = s.y1 - s.y0, ex1 = s.x2 - s.x1, ey1 = s.y2 - s.y1, ex2 = s.x0 - s.x2,
              ey2 = s.y0 - s.y2;
    const I t0 = ehb_edge_inclusive(ex0, ey0, rule) ? 0 : 1, t1 = ehb_edge_inclusive(ex1, ey1, rule) ? 0 : 1,
            t2 = ehb_edge_inclusive(ex2, ey2, rule) ? 0 : 1;
    const float xs = 2.f / (float)W, xo = 1.f / (float)W - 1.f;
    const float ys = 2.f / (float)H, yo = 1.f / (float)H - 1.f;
    const int sx0 = 16 * xlo - bx;
    for (int py = ylo; py <= yhi; py++) {
        const int sy = 16 * py - by;
        I e0 = (I)ex0 * (I)(sy - s.y0) - (I)ey0 * (I)(sx0 - s.x0);
        I e1 = (I)ex1 * (I)(sy - s.y1) - (I)ey1 * (I)(sx0 - s.x1);
        I e2 = (I)ex2 * (I)(sy - s.y2) - (I)ey2 * (I)(sx0 - s.x2);
        for (int px = xlo; px <= xhi; px++) {
            if (e0 >= t0 && e1 >= t1 && e2 >= t2) {
                const float fx = xs * (float)px + xo, fy = ys * (float)py + yo;
                const float zw = ehb_shade_zw(s.c0, s.c1, s.c2, fx, fy);
                const unsigned long long key = ((unsigned long long)ehb_order_key(zw) << 32) | id;
                unsigned long long* dst = pl + (py - ry0) * EHB_RS + (px - rx0);
                if (key < *dst) atomicMin(dst, key);
            }
            e0 -= (I)16 * (I)ey0; e1 -= (I)16 * (I)ey1; e2 -= (I)16 * (I)ey2;
        }
    }
}

// Warp path: lanes stride over the candidate samples of one triangle.
__device__ __forceinline__ void ehb_cover_warp(const EhbTri& s, int xlo, int xhi, int ylo, int yhi,
                                               unsigned long long* pl, int rx0, int ry0, uint32_t id, int H, int W,
                                               int rule, int lane)
{
    const int bx = 8 * W - 8, by = 8 * H - 8;
    const int ex0 = s.x1 - s.x0, ey0 = s.y1 - s.y0, ex1 = s.x2 - s.x1, ey1 = s.y2 - s.y1, ex2 = s.x0 - s.x2,
              ey2 = s.y0 - s.y2;
    const long long t0 = ehb_edge_inclusive(ex0, ey0, rule) ? 0 : 1, t1 = ehb_edge_inclusive(ex1, ey1, rule) ? 0 : 1,
                    t2 = ehb_edge_inclusive(ex2, ey2, rule) ? 0 : 1;
    const float xs = 2.f / (float)W, xo = 1.f / (float)W - 1.f;
    const float ys = 2.f / (float)H, yo = 1.f / (float)H - 1.f;
    const int w = xhi - xlo + 1, n = w * (yhi - ylo + 1);
    for (int i = lane; i < n; i += 32) {
        const int ry = i / w;
        const int px = xlo + (i - ry * w), py = ylo + ry;
        const int sx = 16 * px - bx, sy = 16 * py - by;
        const long long e0 = (long long)ex0 * (sy - s.y0) - (long long)ey0 * (sx - s.x0);
        const long long e1 = (long long)ex1 * (sy - s.y1) - (long long)ey1 * (sx - s.x1);
        const long long e2 = (long long)ex2 * (sy - s.y2) - (long long)ey2 * (sx - s.x2);
        if (e0 >= t0 && e1 >= t1 && e2 >= t2) {
            const float fx = xs * (float)px + xo, fy = ys * (float)py + yo;
            const float zw = ehb_shade_zw(s.c0, s.c1, s.c2, fx, fy);
            const unsigned long long key = ((unsigned long long)ehb_order_key(zw) << 32) | id;
            unsigned long long* dst = pl + (py - ry0) * EHB_RS + (px - rx0);
            if (key < *dst) atomicMin(dst, key);
        }
    }
}

struct EhbRasterSmem {
    float mvp[EHB_MAX_LINKS * 16];
    int links[EHB_MAX_LINKS];     // present links of this tile, ascending
    uint32_t lstart[EHB_MAX_LINKS];
    uint32_t lcnt[EHB_MAX_LINKS];
    int slotOf[EHB_MAX_LINKS];    // link -> plane slot in the current round
    uint32_t big[EHB_BIGQ];
    int nbig;
    int nP;
    int work;
    unsigned int nTiles;
};

// antialiased value of one link at pixel (px,py); lx,ly = position inside the plane
__device__ __forceinline__ float ehb_aa_out(const unsigned long long* pl, const EhbLink& lk, const float* m, int px,
                                            int py, int lx, int ly, int H, int W)
{
    const int idx = ly * EHB_RS + lx;
    const unsigned long long k = pl[idx];
    const bool c = k != EHB_EMPTY;
    const float cf = c ? 1.f : 0.f;
    float o = cf;
    int di;
    if (px < W - 1) {
        const unsigned long long k1 = pl[idx + 1];
        const bool c1 = k1 != EHB_EMPTY;
        if (c1 != c) {
            const float a = ehb_aa_pair(lk, m, (int)(uint32_t)(c ? k : k1), c ? 0 : 1, px, py, 0, H, W, &di);
            if (a > 0.f) o += a * ((c1 ? 1.f : 0.f) - cf);
        }
    }
    if (py < H - 1) {
        const unsigned long long k1 = pl[idx + EHB_RS];
        const bool c1 = k1 != EHB_EMPTY;
        if (c1 != c) {
            const float a = ehb_aa_pair(lk, m, (int)(uint32_t)(c ? k : k1), c ? 0 : 1, px, py, 1, H, W, &di);
            if (a > 0.f) o += a * ((c1 ? 1.f : 0.f) - cf);
        }
    }
    if (px > 0) {
        const unsigned long long k0 = pl[idx - 1];
        const bool c0 = k0 != EHB_EMPTY;
        if (c0 != c) {
            const float a = ehb_aa_pair(lk, m, (int)(uint32_t)(c0 ? k0 : k), c0 ? 0 : 1, px - 1, py, 0, H, W, &di);
            if (!(a > 0.f) && a != 0.f) o += a * (cf - (c0 ? 1.f : 0.f));
        }
    }
    if (py > 0) {
        const unsigned long long k0 = pl[idx - EHB_RS];
        const bool c0 = k0 != EHB_EMPTY;
        if (c0 != c) {
            const float a = ehb_aa_pair(lk, m, (int)(uint32_t)(c0 ? k0 : k), c0 ? 0 : 1, px, py - 1, 1, H, W, &di);
            if (!(a > 0.f) && a != 0.f) o += a * (cf - (c0 ? 1.f : 0.f));
        }
    }
    return o;
}

template <int PMAX>
__global__ void __launch_bounds__(EHB_THREADS) ehb_k_raster(const __grid_constant__ EhbRobot rb,
                                                            const __grid_constant__ EhbParams p)
{
    extern __shared__ __align__(16) unsigned char ehb_smem_raw[];
    unsigned long long* planes = reinterpret_cast<unsigned long long*>(ehb_smem_raw);
    float* sumpl = reinterpret_cast<float*>(planes + PMAX * EHB_NP);
    EhbRasterSmem& sm = *reinterpret_cast<EhbRasterSmem*>(sumpl + ((EHB_NP + 3) & ~3));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = p.H, W = p.W;
    const bool perLink = p.Lk != 1;

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
            if (lane == 0) { sm.nP = __popc(b); sm.nbig = 0; }
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
            if (tid == 0) sm.nbig = 0;
            __syncthreads();
            const uint32_t lo = sm.lstart[k0];
            const uint32_t n = sm.lstart[k1 - 1] + sm.lcnt[k1 - 1] - lo;
            for (uint32_t i = tid; i < n; i += EHB_THREADS) {
                const uint32_t e = __ldg(p.pairs + lo + i);
                const int l = e >> EHB_LINK_SHIFT;
                const int f = e & EHB_FACE_MASK;
                EhbTri s;
                if (ehb_tri_setup(rb.link[l], sm.mvp + 16 * l, f, H, W, s)) continue;
                const int xlo = max(s.pxlo, rx0), xhi = min(s.pxhi, rx1);
                const int ylo = max(s.pylo, ry0), yhi = min(s.pyhi, ry1);
                if (xlo > xhi || ylo > yhi) continue;
                if ((xhi - xlo + 1) * (yhi - ylo + 1) > EHB_BIG_SAMPLES) {
                    const int k = atomicAdd(&sm.nbig, 1);
                    if (k < EHB_BIGQ) { sm.big[k] = e; continue; }
                }
                unsigned long long* pl = planes + sm.slotOf[l] * EHB_NP;
                const uint32_t id = perLink ? (uint32_t)f : (uint32_t)(rb.foff[l] + f);
                const int ext = max(max(s.x0, max(s.x1, s.x2)) - min(s.x0, min(s.x1, s.x2)),
                                    max(s.y0, max(s.y1, s.y2)) - min(s.y0, min(s.y1, s.y2)));
                if (ext < 32768) ehb_cover_rows<int>(s, xlo, xhi, ylo, yhi, pl, rx0, ry0, id, H, W, p.rule);
                else ehb_cover_rows<long long>(s, xlo, xhi, ylo, yhi, pl, rx0, ry0, id, H, W, p.rule);
            }
            __syncthreads();
            const int nb = min(sm.nbig, EHB_BIGQ);
            for (int k = warp; k < nb; k += EHB_THREADS / 32) {
                const uint32_t e = sm.big[k];
                const int l = e >> EHB_LINK_SHIFT;
                const int f = e & EHB_FACE_MASK;
                EhbTri s;
                if (ehb_tri_setup(rb.link[l], sm.mvp + 16 * l, f, H, W, s)) continue;
                const int xlo = max(s.pxlo, rx0), xhi = min(s.pxhi, rx1);
                const int ylo = max(s.pylo, ry0), yhi = min(s.pyhi, ry1);
                unsigned long long* pl = planes + sm.slotOf[l] * EHB_NP;
                const uint32_t id = perLink ? (uint32_t)f : (uint32_t)(rb.foff[l] + f);
                ehb_cover_warp(s, xlo, xhi, ylo, yhi, pl, rx0, ry0, id, H, W, p.rule, lane);
            }
            __syncthreads();
        };

        // ---- antialias backward of the links resident in the planes ----------------------------------------
        auto backward_round = [&](int r) {
            const int k0 = r * PMAX, k1 = min(nP, k0 + PMAX);
            for (int k = k0; k < k1; k++) {
                const int l = sm.links[k];
                const EhbLink& lk = rb.link[l];
                const float* m = sm.mvp + 16 * l;
                const unsigned long long* pl = planes + (k - k0) * EHB_NP;
                double acc[12];
#pragma unroll
                for (int i = 0; i < 12; i++) acc[i] = 0.0;
                bool had = false;
                for (int i = tid; i < EHB_T * EHB_T; i += EHB_THREADS) {
                    const int px = x0 + (i & 31), py = y0 + (i >> 5);
                    if (px >= W || py >= H) continue;
                    const int idx = (py - ry0) * EHB_RS + (px - rx0);
                    const unsigned long long ka = pl[idx];
                    const bool c0 = ka != EHB_EMPTY;
#pragma unroll 1
                    for (int d = 0; d < 2; d++) {
                        if (d == 0 ? (px >= W - 1) : (py >= H - 1)) continue;
                        const int idx1 = idx + (d ? EHB_RS : 1);
                        const unsigned long long kb = pl[idx1];
                        const bool c1 = kb != EHB_EMPTY;
                        if (c0 == c1) continue;
                        const int side = c0 ? 0 : 1;
                        const int t = (int)(uint32_t)(c0 ? ka : kb);
                        int di;
                        const float al = ehb_aa_pair(lk, m, t, side, px, py, d, H, W, &di);
                        if (al == 0.f) continue;
                        const float g = sumpl[al > 0.f ? idx : idx1];
                        const float dd = g * ((c1 ? 1.f : 0.f) - (c0 ? 1.f : 0.f));
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
                            for (int c = 0; c < 4; c++)
                                acc[4 * rr + c] += (double)g1[rr] * ha[c] + (double)g2[rr] * hb[c];
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
        };

        // ======================================= forward =====================================================
        if (needAA)
            for (int i = tid; i < EHB_NP; i += EHB_THREADS) sumpl[i] = 0.f;
        for (int r = 0; r < nR; r++) {
            raster_round(r);
            if (needAA) {
                const int k0 = r * PMAX, k1 = min(nP, k0 + PMAX);
                for (int i = tid; i < ow * ow; i += EHB_THREADS) {
                    const int qx = i % ow, qy = i / ow;
                    const int px = x0 + qx, py = y0 + qy;
                    if (px >= W || py >= H) continue;
                    const int lx = px - rx0, ly = py - ry0;
                    float a = sumpl[ly * EHB_RS + lx];
                    for (int k = k0; k < k1; k++) {
                        const int l = sm.links[k];
                        const float o = ehb_aa_out(planes + (k - k0) * EHB_NP, rb.link[l], sm.mvp + 16 * l, px, py, lx,
                                                   ly, H, W);
                        a = a + o;  // links are added in link order (rb_solver.py:68); absent links add exactly 0
                    }
                    sumpl[ly * EHB_RS + lx] = a;
                }
                __syncthreads();
            }
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
                const int qx = i % ow, qy = i / ow;
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
                if (nR > 1) __syncthreads();
            }
        }
        __syncthreads();
    }
}
