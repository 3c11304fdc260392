// ehb_tiles.cuh -- the image-space half of a pass: antialias, compose, loss, backward.
//
// The reference antialiases every link on its own, sums the per-link masks and clamps (rb_solver.py:62-68), and its
// backward scatters one gradient per silhouette pixel pair (dr.antialias).  On the chip this work is a
// few thousand small, latency-bound jobs per pass; what matters is that ALL of them are in flight at once and that no
// job waits for another.  So the stage is three flat kernels over three work lists instead of one CTA per tile that
// walks its links, its phases and its pairs in sequence:
//   (jobs)     a job = one (tile, link) window that some triangle of the link reaches into; listed by spare CTAs of
//              k_raster_big from the touch bitmap, the jobs of a tile are consecutive and in link order
//   k_windows  one WARP per job, no block barrier: 35x35 window of the link's depth plane -> row coverage masks by
//              ballot -> silhouette pairs by XOR of neighbouring masks -> blend weights (one pair per lane) -> the
//              link's antialiased mask of the tile's out region (33 x 33, L2-resident scratch) + the pair entries
//   k_compose  one warp per tile: S = min(sum of its jobs' masks in link order, 1), mask write, (S - ref)^2,
//              g = dL/dsum into the tile's gradient window
//   k_pairgrad one THREAD per pair entry: analytic vertex gradients contracted with [x y z 1] on the fly, warp-shuffle
//              reduction per (item, link), fp64 atomicAdd into d loss / d mvp
#pragma once
#include "ehb_kernels.cuh"

#define EHB_MROWS (EHB_T + 1)            // out region of a tile: interior + one row / column on the high side
#define EHB_MW 36                        // row pitch (floats) of a job mask / gradient window: 16-byte aligned rows
#define EHB_MSZ (EHB_MROWS * EHB_MW)
#define EHB_NSEG (EHB_T * 4)             // out region: 32 rows x 4 eight-pixel segments ...
#define EHB_NSEG_EXT (4 + EHB_T + 1)     // ... + row 32 (4 segments) + column 32 (33 single pixels) when the backward follows
#ifndef EHB_WWARPS
#define EHB_WWARPS 4                     // warps (jobs in flight) per k_windows CTA
#endif
#define EHB_WPAIRS 256                   // pairs of a job whose packed word / blend weight stay in shared memory

// Out-region work item s -> (row qy, first column qx0, pixels n): eight consecutive pixels of one row, then the extra
// row / column that exist when the backward follows in the same pass.
__device__ __forceinline__ void ehb_out_segment(int s, int& qy, int& qx0, int& n)
{
    if (s < EHB_NSEG) { qy = s >> 2; qx0 = (s & 3) * 8; n = 8; }
    else if (s < EHB_NSEG + 4) { qy = EHB_T; qx0 = (s - EHB_NSEG) * 8; n = 8; }
    else { qy = s - EHB_NSEG - 4; qx0 = EHB_T; n = 1; }   // column 32, corner included
}

__device__ __forceinline__ unsigned long long ehb_bits(int lo, int hi)   // bits lo..hi (inclusive), empty if lo > hi
{
    lo = max(lo, 0); hi = min(hi, 63);
    if (lo > hi) return 0ull;
    const unsigned long long up = hi >= 63 ? ~0ull : ((1ull << (hi + 1)) - 1ull);
    return up & ~((1ull << lo) - 1ull);
}

__device__ __forceinline__ uint32_t ehb_list_at(const EhbParams& p, unsigned e, unsigned nHeavy)
{
    return p.tileList[e < nHeavy ? e : (unsigned)(p.items * p.ntiles) - 1u - (e - nHeavy)];
}

// ------------------------------------------------------------------------------------------------ job list
// One lane per listed tile: its touch bits become consecutive jobs (link order); one queue atomic per warp.
__device__ __forceinline__ void ehb_build_jobs(const EhbParams& p, int firstWarp, int nWarps, int lane)
{
    const unsigned nHeavy = p.ctr->nTiles, nEntries = nHeavy + p.ctr->nLight;
    for (unsigned e0 = (unsigned)firstWarp * 32u; e0 < nEntries; e0 += (unsigned)nWarps * 32u) {
        const unsigned e = e0 + (unsigned)lane;
        uint32_t wid = 0, bits = 0;
        if (e < nEntries) {
            wid = ehb_list_at(p, e, nHeavy);
            bits = p.touch[wid] & (p.L >= 32 ? 0xFFFFFFFFu : ((1u << p.L) - 1u));
        }
        const int n = __popc(bits);
        int inc = n;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        const int tot = __shfl_sync(0xffffffffu, inc, 31);
        unsigned base = 0;
        if (lane == 0 && tot > 0) base = atomicAdd(&p.ctr->nJobs, (unsigned)tot);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (e < nEntries) {
            unsigned j = base + (unsigned)(inc - n);
            p.tileEnt[e] = make_uint4(wid, bits, j, 0u);
            const int item = (int)(wid / (uint32_t)p.ntiles), tile = (int)(wid - (uint32_t)item * (uint32_t)p.ntiles);
            if (n && j + (unsigned)n > (unsigned)p.jobCap) atomicOr(&p.ctr->flags, 1u);   // job list too small: grow and rerun
            while (bits) {
                const int l = __ffs(bits) - 1;
                bits &= bits - 1;
                const EhbPlane pl = p.plane[(size_t)item * p.L + l];
                EhbJob jb;
                jb.item = item; jb.tile = tile; jb.link = l; jb.entry = (int)e;
                jb.x0 = pl.x0; jb.y0 = pl.y0; jb.w = pl.w; jb.h = pl.h; jb.off = pl.off; jb.pad = 0;
                if (j < (unsigned)p.jobCap) p.jobs[j] = jb;
                j++;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ k_windows
struct __align__(16) EhbWarpSm {
    unsigned long long cov[EHB_RS + 1];
    // triangle id of the nearest sample of the link (0xFFFFFFFF = not covered); once the blend weights are known the
    // same storage holds the job's antialiased mask (33 rows x 36 floats)
    uint32_t plane[EHB_NP + 3];
    float alpha[EHB_WPAIRS];
    unsigned short pk[EHB_WPAIRS];
};
static_assert(EHB_NP + 3 >= EHB_MSZ, "the mask of a job reuses the window's storage");

__global__ void __launch_bounds__(EHB_WWARPS * 32) ehb_k_windows(const __grid_constant__ EhbRobot rb,
                                                                const __grid_constant__ EhbParams p)
{
    ehb_pdl_enter();
    __shared__ EhbWarpSm s_w[EHB_WWARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    EhbWarpSm& ws = s_w[warp];
    const int H = p.H, W = p.W, hlo = p.hlo;
    const bool needAA = p.mode == EHB_MODE_FUSED || p.mode == EHB_MODE_AA_FWD;
    const int oext = (p.mode == EHB_MODE_FUSED && p.do_bwd) ? 1 : 0;
    const int ow = EHB_T + oext;
    // the first job record is fetched together with the job counter (one L2 round trip instead of two)
    const unsigned jFirst = blockIdx.x * EHB_WWARPS + warp;
    const EhbJob jbFirst = p.jobs[min(jFirst, (unsigned)p.jobCap - 1u)];
    const unsigned nJobs = min(p.ctr->nJobs, (unsigned)p.jobCap);
    for (unsigned j = jFirst; j < nJobs; j += gridDim.x * EHB_WWARPS) {
        const EhbJob jb = j == jFirst ? jbFirst : p.jobs[j];
        const int item = jb.item, l = jb.link;
        const int tx = jb.tile % p.ntx, ty = jb.tile / p.ntx;
        const int rx0 = tx * EHB_T - hlo, ry0 = ty * EHB_T - hlo;
        const int wcols = EHB_T + hlo + p.hhi, wrows = wcols;   // window size: 34 or 35 (33 for the operator backward)
        EhbPlane pl;
        pl.x0 = jb.x0; pl.y0 = jb.y0; pl.w = jb.w; pl.h = jb.h; pl.off = jb.off;
        const EhbLink& lk = rb.link[l];
        const float4* vc = p.vclip + (size_t)item * p.Vtot + rb.voff[l];
        __syncwarp();   // the previous job of this warp is done with ws
        // ================ window of the link's plane -> shared memory, row coverage masks by ballot ================
        unsigned long long myCov = 0ull;   // lane r: coverage mask of window row r (rows 32.. : lanes 0..2, second word)
        unsigned long long myCov2 = 0ull;
        {
            const int cx = rx0 + lane - pl.x0;
            const bool colOk = pl.w > 0 && cx >= 0 && cx < pl.w && lane < wcols;
            const unsigned long long* base = p.pool + pl.off + cx;
            const int pyBase = ry0 - pl.y0;
            // columns 0..31 of every row: the loads of NRC rows are issued before the first ballot consumes one
            constexpr int NRC = 12;
#pragma unroll
            for (int rbase = 0; rbase < 36; rbase += NRC) {
                unsigned long long v[NRC];
#pragma unroll
                for (int k = 0; k < NRC; k++) {
                    const int r = rbase + k, py = pyBase + r;
                    v[k] = EHB_EMPTY;
                    if (r < EHB_RS && colOk && (unsigned)py < (unsigned)pl.h && r < wrows) v[k] = base[(long long)py * pl.w];
                }
#pragma unroll
                for (int k = 0; k < NRC; k++) {
                    const int r = rbase + k;
                    if (r < EHB_RS) {
                        ws.plane[r * EHB_RS + lane] = (uint32_t)v[k];   // low word = triangle id (all ones when empty)
                        const unsigned b = __ballot_sync(0xffffffffu, v[k] != EHB_EMPTY);
                        if (r < 32) { if (lane == r) myCov = (unsigned long long)b; }
                        else if (lane == r - 32) myCov2 = (unsigned long long)b;
                    }
                }
            }
            // columns 32..34: the 105 elements as one flat list (4 loads instead of 35 mostly idle ones)
            unsigned bx[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int i = k * 32 + lane;
                const int r = i / 3, c = 32 + (i - r * 3);
                unsigned long long v = EHB_EMPTY;
                if (i < 3 * EHB_RS) {
                    const int py = pyBase + r, cxx = rx0 + c - pl.x0;
                    if (pl.w > 0 && cxx >= 0 && cxx < pl.w && c < wcols && (unsigned)py < (unsigned)pl.h && r < wrows)
                        v = p.pool[pl.off + (long long)py * pl.w + cxx];
                    ws.plane[r * EHB_RS + c] = (uint32_t)v;
                }
                bx[k] = __ballot_sync(0xffffffffu, v != EHB_EMPTY);
            }
            // row r owns bits 3r .. 3r+2 of the 128-bit string bx[3]:bx[2]:bx[1]:bx[0]
            {
                const unsigned long long lo = (unsigned long long)bx[0] | ((unsigned long long)bx[1] << 32);
                const unsigned long long hi = (unsigned long long)bx[2] | ((unsigned long long)bx[3] << 32);
                auto three = [&](int r) -> unsigned long long {
                    const int s = 3 * r;
                    unsigned long long w = s < 64 ? (lo >> s) : (hi >> (s - 64));
                    if (s < 64 && s > 61) w |= hi << (64 - s);
                    return w & 7ull;
                };
                myCov |= three(lane) << 32;
                if (lane < EHB_RS - 32) myCov2 |= three(lane + 32) << 32;
            }
            ws.cov[lane] = myCov;
            if (lane < EHB_RS - 32) ws.cov[lane + 32] = myCov2;
            if (lane == EHB_RS - 32) ws.cov[EHB_RS] = 0ull;
        }
        __syncwarp();
        // ================================ silhouette pairs: lane = window row ====================================
        int nPairs;
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
                const unsigned long long cm = h ? myCov2 : myCov, cu = ws.cov[r + 1];
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
        nPairs = tot0 + tot1;
        const int o0 = inc - cnt[0], o1 = tot0 + inc1 - cnt[1];
        // The list of this job in the global pair pool.  The ticket is drawn now and used after the blend weights are
        // known (they only need shared memory), so the atomic's round trip hides behind the weights' own loads; only a
        // job with more pairs than the shared-memory arrays hold waits for it right away.
        unsigned ticket = 0;
        if (lane == 0 && nPairs > 0) ticket = atomicAdd(&p.ctr->pairCursor, (unsigned)nPairs);
        EhbPair* mine = nullptr;
        bool resolved = false;
        auto resolve = [&]() {
            const unsigned pairOff = __shfl_sync(0xffffffffu, ticket, 0);
            resolved = true;
            mine = p.pairs + pairOff;
            if ((unsigned long long)pairOff + (unsigned long long)nPairs > (unsigned long long)p.pairCap) {
                if (lane == 0) atomicOr(&p.ctr->flags, 1u);   // pair pool too small: results invalid, grow and rerun
                mine = nullptr;
            }
        };
        if (nPairs > EHB_WPAIRS) {
            resolve();
            if (!mine) nPairs = 0;
        }
        if (nPairs > 0) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                int o = h ? o1 : o0;
                unsigned long long hxm = hm[h], vym = vm[h];
                const uint32_t rowBase = (uint32_t)(lane + 32 * h) * EHB_RS;
                while (hxm) {
                    const int b = __ffsll((long long)hxm) - 1;
                    hxm &= hxm - 1;
                    const uint32_t pk = (rowBase + b) | (((om[h] >> b) & 1ull) ? (1u << 12) : 0u);
                    if (o < EHB_WPAIRS) ws.pk[o] = (unsigned short)pk; else mine[o].packed = pk;
                    o++;
                }
                while (vym) {
                    const int b = __ffsll((long long)vym) - 1;
                    vym &= vym - 1;
                    const uint32_t pk = (rowBase + b) | (1u << 11) | (((om[h] >> b) & 1ull) ? (1u << 12) : 0u);
                    if (o < EHB_WPAIRS) ws.pk[o] = (unsigned short)pk; else mine[o].packed = pk;
                    o++;
                }
            }
        }
        __syncwarp();
        // ================================ blend weights, one pair per lane =====================================
        auto store_entry = [&](int i, uint32_t pk2, uint32_t t, float al) {
            uint4* e = reinterpret_cast<uint4*>(mine + i);
            e[0] = make_uint4(pk2, t, __float_as_uint(al), j);
            e[1] = make_uint4((uint32_t)item, (uint32_t)jb.tile, (uint32_t)l, (uint32_t)jb.entry);
        };
        auto tri_of = [&](uint32_t pk, int& side) -> uint32_t {
            const int idx = pk & 2047, d = (pk >> 11) & 1;
            const uint32_t ka = ws.plane[idx], kb = ws.plane[idx + (d ? EHB_RS : 1)];
            side = ka != 0xFFFFFFFFu ? 0 : 1;
            return side ? kb : ka;
        };
        for (int i = lane; i < nPairs; i += 32) {
            const uint32_t pk = i < EHB_WPAIRS ? (uint32_t)ws.pk[i] : __ldcg(&mine[i].packed);
            const int idx = pk & 2047, d = (pk >> 11) & 1;
            const int ly = idx / EHB_RS, lx = idx - ly * EHB_RS;
            int side, di;
            const uint32_t t = tri_of(pk, side);
            const float al = ehb_aa_pair(lk, vc, (int)t, side, rx0 + lx, ry0 + ly, d, H, W, &di);
            const uint32_t pk2 = pk | ((uint32_t)side << 13) | ((uint32_t)di << 14);
            if (resolved) store_entry(i, pk2, t, al);
            if (i < EHB_WPAIRS) { ws.alpha[i] = al; ws.pk[i] = (unsigned short)(pk2 & 0xFFFFu); }
        }
        if (!resolved && nPairs > 0) {   // nPairs <= EHB_WPAIRS: everything is in shared memory; (each lane re-reads its own entries)
            resolve();
            if (mine)
                for (int i = lane; i < nPairs; i += 32) {
                    const uint32_t pk2 = ws.pk[i];
                    int side;
                    const uint32_t t = tri_of(pk2, side);
                    store_entry(i, pk2, t, ws.alpha[i]);
                }
        }
        if (!needAA) continue;
        __syncwarp();
        // ================= the link's antialiased mask of the out region: colour, then the pair contributions ============
        // (the window's triangle ids are dead: their storage becomes the mask)
        float* ot = reinterpret_cast<float*>(ws.plane);
        for (int sg = lane; sg < EHB_NSEG + (oext ? EHB_NSEG_EXT : 0); sg += 32) {
            int qy, qx0, n;
            ehb_out_segment(sg, qy, qx0, n);
            const uint32_t c8 = (uint32_t)(ws.cov[hlo + qy] >> (hlo + qx0));
            float* dst = ot + qy * EHB_MW + qx0;
            if (n == 8) {
                reinterpret_cast<float4*>(dst)[0] = make_float4((c8 & 1u) ? 1.f : 0.f, (c8 & 2u) ? 1.f : 0.f, (c8 & 4u) ? 1.f : 0.f, (c8 & 8u) ? 1.f : 0.f);
                reinterpret_cast<float4*>(dst)[1] = make_float4((c8 & 16u) ? 1.f : 0.f, (c8 & 32u) ? 1.f : 0.f, (c8 & 64u) ? 1.f : 0.f, (c8 & 128u) ? 1.f : 0.f);
            } else dst[0] = (c8 & 1u) ? 1.f : 0.f;
        }
        // A pixel receives at most one contribution of each kind, and the reference adds them in the order
        // pair(p,p+x), pair(p,p+y), pair(p-x,p), pair(p-y,p): four sweeps over the pair list, one kind each
        // (receiver = p0 when alpha > 0, p1 otherwise; the contribution is alpha * (colour[p1] - colour[p0])).
#pragma unroll 1
        for (int kind = 0; kind < 4; kind++) {
            __syncwarp();
            for (int i = lane; i < nPairs; i += 32) {
                uint32_t pk; float al;
                if (i < EHB_WPAIRS) { pk = ws.pk[i]; al = ws.alpha[i]; }
                else { const uint4 e = __ldcg(reinterpret_cast<const uint4*>(mine + i)); pk = e.x; al = __uint_as_float(e.z); }
                const int d = (pk >> 11) & 1;
                const bool pos = al > 0.f;
                if (al == 0.f || d != (kind & 1) || pos != (kind < 2)) continue;
                const int idx = pk & 2047, side = (pk >> 13) & 1;
                const int ridx = pos ? idx : idx + (d ? EHB_RS : 1);
                const int ry = ridx / EHB_RS, rxw = ridx - ry * EHB_RS;
                const int qy = ry - hlo, qx = rxw - hlo;
                if (qy < 0 || qx < 0 || qy >= ow || qx >= ow) continue;
                const float delta = side ? 1.f : -1.f;   // colour[p1] - colour[p0]: p1 is the covered one when side = 1
                ot[qy * EHB_MW + qx] += al * delta;
            }
        }
        __syncwarp();
        {   // mask -> the job's slot (coalesced 16-byte stores)
            float4* dst = reinterpret_cast<float4*>(p.maskBuf + (size_t)j * EHB_MSZ);
            const float4* src = reinterpret_cast<const float4*>(ot);
            const int n4 = (EHB_T + oext) * (EHB_MW / 4);
            for (int i = lane; i < n4; i += 32) dst[i] = src[i];
        }
    }
}

// ------------------------------------------------------------------------------------------------ k_compose
// One 128-thread CTA per listed tile, one out-region segment (8 pixels) per thread: sum of the tile's job masks in link
// order (rb_solver.py:68), clamp, mask write, loss, dL/dsum.  Every load of a thread is issued before the first is used.
#define EHB_CTHREADS 128
__global__ void __launch_bounds__(EHB_CTHREADS) ehb_k_compose(const __grid_constant__ EhbParams p)
{
    ehb_pdl_enter();
    const int tid = threadIdx.x, lane = tid & 31;
    const int H = p.H, W = p.W;
    const bool fused = p.mode == EHB_MODE_FUSED;
    const int oext = (fused && p.do_bwd) ? 1 : 0;
    const bool haveRef = p.ref != nullptr || p.ref_u8 != nullptr;
    const unsigned nHeavy = p.ctr->nTiles, nEntries = nHeavy + p.ctr->nLight;
    const bool vecOut = p.masks != nullptr && (W & 3) == 0 && (((uintptr_t)p.masks) & 15) == 0;
    const bool vecRef = p.ref != nullptr && (W & 3) == 0 && (((uintptr_t)p.ref) & 15) == 0;
    const bool vecRef8 = p.ref_u8 != nullptr && (W & 7) == 0 && (((uintptr_t)p.ref_u8) & 7) == 0;
    const unsigned eFirst = blockIdx.x;
    const uint4 teFirst = p.tileEnt[min(eFirst, (unsigned)(p.items * p.ntiles) - 1u)];   // fetched together with the counters
    for (unsigned e = eFirst; e < nEntries; e += gridDim.x) {
        const uint4 te = e == eFirst ? teFirst : p.tileEnt[e];   // {wid, link bits, first job, -}, written by the job builder
        const uint32_t wid = te.x;
        const int item = (int)(wid / (uint32_t)p.ntiles), tile = (int)(wid - (uint32_t)item * (uint32_t)p.ntiles);
        const int x0 = (tile % p.ntx) * EHB_T, y0 = (tile / p.ntx) * EHB_T;
        const size_t ibase = (size_t)item * H * W;
        float* gwin = p.gBuf + (size_t)e * EHB_MSZ;
        if (p.mode == EHB_MODE_AA_BWD) {   // g = dL/dmask comes from the caller
            for (int i = tid; i < EHB_MROWS * EHB_MROWS; i += EHB_CTHREADS) {
                const int qy = i / EHB_MROWS, qx = i - qy * EHB_MROWS;
                const int px = x0 + qx, py = y0 + qy;
                gwin[qy * EHB_MW + qx] = (px < W && py < H) ? __ldg(p.dy + ibase + (size_t)(H - 1 - py) * W + px) : 0.f;
            }
            continue;
        }
        const uint32_t bits = te.y;
        const unsigned j0 = te.z;
        const int nP = min(__popc(bits), (int)max(0ll, (long long)p.jobCap - (long long)j0));   // (overflow: flagged, rerun)
        const float* m0 = p.maskBuf + (size_t)j0 * EHB_MSZ;
        double lacc = 0.0;
        for (int sg = tid; sg < EHB_NSEG + (oext ? EHB_NSEG_EXT : 0); sg += EHB_CTHREADS) {
            int qy, qx0, n;
            ehb_out_segment(sg, qy, qx0, n);
            const int py = y0 + qy;
            if (py >= H) continue;
            const size_t orow = ibase + (size_t)(H - 1 - py) * W;
            const int px0 = x0 + qx0;
            // reference of the segment (issued before the masks are consumed)
            float rf[8];
#pragma unroll
            for (int i = 0; i < 8; i++) rf[i] = 0.f;
            if (fused && haveRef) {
                if (n == 8 && px0 + 7 < W && vecRef) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(p.ref + orow + px0));
                    const float4 b = __ldg(reinterpret_cast<const float4*>(p.ref + orow + px0) + 1);
                    rf[0] = a.x; rf[1] = a.y; rf[2] = a.z; rf[3] = a.w; rf[4] = b.x; rf[5] = b.y; rf[6] = b.z; rf[7] = b.w;
                } else if (n == 8 && px0 + 7 < W && vecRef8) {
                    const uint2 a = __ldg(reinterpret_cast<const uint2*>(p.ref_u8 + orow + px0));
#pragma unroll
                    for (int i = 0; i < 8; i++) rf[i] = (((i < 4 ? a.x : a.y) >> (8 * (i & 3))) & 255u) ? 1.f : 0.f;
                } else {
                    for (int i = 0; i < n; i++)
                        if (px0 + i < W) rf[i] = p.ref ? __ldg(p.ref + orow + px0 + i) : (__ldg(p.ref_u8 + orow + px0 + i) ? 1.f : 0.f);
                }
            }
            float s[8];
#pragma unroll
            for (int i = 0; i < 8; i++) s[i] = 0.f;
            const float* src = m0 + qy * EHB_MW + qx0;
            if (n == 8) {
                int k = 0;
                for (; k + 1 < nP; k += 2) {   // links are added in link order; two jobs' loads in flight at a time
                    const float4 a0 = __ldcg(reinterpret_cast<const float4*>(src + (size_t)k * EHB_MSZ));
                    const float4 b0 = __ldcg(reinterpret_cast<const float4*>(src + (size_t)k * EHB_MSZ) + 1);
                    const float4 a1 = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(k + 1) * EHB_MSZ));
                    const float4 b1 = __ldcg(reinterpret_cast<const float4*>(src + (size_t)(k + 1) * EHB_MSZ) + 1);
                    s[0] = (s[0] + a0.x) + a1.x; s[1] = (s[1] + a0.y) + a1.y; s[2] = (s[2] + a0.z) + a1.z; s[3] = (s[3] + a0.w) + a1.w;
                    s[4] = (s[4] + b0.x) + b1.x; s[5] = (s[5] + b0.y) + b1.y; s[6] = (s[6] + b0.z) + b1.z; s[7] = (s[7] + b0.w) + b1.w;
                }
                if (k < nP) {
                    const float4 a = __ldcg(reinterpret_cast<const float4*>(src + (size_t)k * EHB_MSZ));
                    const float4 b = __ldcg(reinterpret_cast<const float4*>(src + (size_t)k * EHB_MSZ) + 1);
                    s[0] = s[0] + a.x; s[1] = s[1] + a.y; s[2] = s[2] + a.z; s[3] = s[3] + a.w;
                    s[4] = s[4] + b.x; s[5] = s[5] + b.y; s[6] = s[6] + b.z; s[7] = s[7] + b.w;
                }
            } else {
                for (int k = 0; k < nP; k++) s[0] = s[0] + __ldcg(src + (size_t)k * EHB_MSZ);
            }
            float Sv[8], gv[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                Sv[i] = 0.f; gv[i] = 0.f;
                if (i >= n || px0 + i >= W) continue;
                const float S = (p.clamp && s[i] > 1.f) ? 1.f : s[i];
                Sv[i] = S;
                if (fused && haveRef) {
                    const float diff = S - rf[i];
                    if (qx0 + i < EHB_T && qy < EHB_T) lacc += (double)(diff * diff);
                    gv[i] = (!p.clamp || s[i] <= 1.f) ? (2.f * diff) * p.invB : 0.f;
                }
            }
            if (oext) {
                float* gd = gwin + qy * EHB_MW + qx0;
                if (n == 8) {
                    reinterpret_cast<float4*>(gd)[0] = make_float4(gv[0], gv[1], gv[2], gv[3]);
                    reinterpret_cast<float4*>(gd)[1] = make_float4(gv[4], gv[5], gv[6], gv[7]);
                } else gd[0] = gv[0];
            }
            if (p.masks && qy < EHB_T && qx0 < EHB_T) {
                if (vecOut && px0 + 7 < W) {
                    float4* dst = reinterpret_cast<float4*>(p.masks + orow + px0);
                    dst[0] = make_float4(Sv[0], Sv[1], Sv[2], Sv[3]);
                    dst[1] = make_float4(Sv[4], Sv[5], Sv[6], Sv[7]);
                } else {
                    for (int i = 0; i < n; i++)
                        if (px0 + i < W) p.masks[orow + px0 + i] = Sv[i];
                }
            }
        }
        if (fused && haveRef && p.loss) {
            lacc = ehb_warp_sum(lacc);
            if (lane == 0 && lacc != 0.0) atomicAdd(&p.loss[item], lacc);
        }
    }
}

// ------------------------------------------------------------------------------------------------ k_pairgrad
// One thread per pair entry.  Owned pairs (p0 inside the tile's interior) with a non-zero weight and a non-zero upstream
// gradient at their receiving pixel contribute; everything of a warp that belongs to one (item, link) is reduced by
// shuffles before the fp64 atomics.
__global__ void __launch_bounds__(256) ehb_k_pairgrad(const __grid_constant__ EhbRobot rb, const __grid_constant__ EhbParams p)
{
    ehb_pdl_enter();
    const int lane = threadIdx.x & 31;
    const unsigned nPairs = min(p.ctr->pairCursor, (unsigned)p.pairCap);
    const unsigned stride = gridDim.x * blockDim.x;
    const unsigned nIter = (nPairs + stride - 1) / stride;
    const int hlo = p.hlo;
    for (unsigned it = 0; it < nIter; it++) {
        const unsigned i = it * stride + blockIdx.x * blockDim.x + threadIdx.x;
        double acc[12];
#pragma unroll
        for (int k = 0; k < 12; k++) acc[k] = 0.0;
        int key = -1;
        if (i < nPairs) {
            const uint4 e = __ldcg(reinterpret_cast<const uint4*>(p.pairs + i));
            const uint4 jq = __ldcg(reinterpret_cast<const uint4*>(p.pairs + i) + 1);   // item, tile, link, entry
            const float al = __uint_as_float(e.z);
            // (after a pool overflow -- flagged, the pass is rerun -- the tail of the pool may hold stale entries: never
            // follow an index that is out of range)
            const bool sane = jq.x < (unsigned)p.items && jq.y < (unsigned)p.ntiles && jq.z < (unsigned)p.L &&
                              jq.w < (unsigned)(p.items * p.ntiles) && e.y < (unsigned)rb.link[min(jq.z, (unsigned)p.L - 1u)].F;
            if (sane && (e.x & (1u << 12)) && al != 0.f) {
                const int jitem = (int)jq.x, jtile = (int)jq.y, jlink = (int)jq.z, jentry = (int)jq.w;
                const int idx = e.x & 2047, d = (e.x >> 11) & 1, side = (e.x >> 13) & 1, di = (e.x >> 14) & 3;
                const int idx1 = idx + (d ? EHB_RS : 1);
                const int ridx = al > 0.f ? idx : idx1;
                const int ry = ridx / EHB_RS, rxw = ridx - ry * EHB_RS;
                const float g = __ldcg(p.gBuf + (size_t)jentry * EHB_MSZ + (ry - hlo) * EHB_MW + (rxw - hlo));
                const float dd = g * (side ? 1.f : -1.f);   // g * (c1 - c0)
                if (dd != 0.f) {
                    const int ly = idx / EHB_RS, lx = idx - ly * EHB_RS;
                    const int rx0 = (jtile % p.ntx) * EHB_T - hlo, ry0 = (jtile / p.ntx) * EHB_T - hlo;
                    const EhbLink& lk = rb.link[jlink];
                    const float4* vc = p.vclip + (size_t)jitem * p.Vtot + rb.voff[jlink];
                    int vi1, vi2;
                    float g1[3], g2[3];
                    ehb_aa_pair_grad(lk, vc, (int)e.y, side, di, al, dd, rx0 + lx, ry0 + ly, d, p.H, p.W, &vi1, &vi2, g1, g2);
                    const float4 va = __ldg(lk.verts + vi1), vb = __ldg(lk.verts + vi2);
                    const double ha[4] = {(double)va.x, (double)va.y, (double)va.z, 1.0};
                    const double hb[4] = {(double)vb.x, (double)vb.y, (double)vb.z, 1.0};
#pragma unroll
                    for (int rr = 0; rr < 3; rr++)
#pragma unroll
                        for (int c = 0; c < 4; c++) acc[4 * rr + c] = (double)g1[rr] * ha[c] + (double)g2[rr] * hb[c];
                    key = jitem * p.L + jlink;
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
        }
        unsigned todo = __ballot_sync(0xffffffffu, key >= 0);
        while (todo) {
            const int leader = __ffs(todo) - 1;
            const int k0 = __shfl_sync(0xffffffffu, key, leader);
            const bool mine = key == k0;
            todo &= ~__ballot_sync(0xffffffffu, mine);
            double* dst = p.gmvp + (size_t)k0 * 16;
#pragma unroll
            for (int k = 0; k < 12; k++) {
                const double v = ehb_warp_sum(mine ? acc[k] : 0.0);
                // rows x (0), y (1), w (3) of d loss / d mvp; the z row carries no gradient
                if (lane == 0 && v != 0.0) atomicAdd(dst + (k < 8 ? k : k + 4), v);
            }
        }
    }
}
