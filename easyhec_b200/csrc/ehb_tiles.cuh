// ehb_tiles.cuh -- the image-space half of a pass: antialias, compose, loss, backward -- ONE kernel.
//
// The reference antialiases every link on its own, sums the per-link masks and clamps (rb_solver.py:62-68), and its
// backward scatters one gradient per silhouette pixel pair (dr.antialias).  Here one CTA owns one listed
// 32x32 tile and keeps everything of it in shared memory, so that no intermediate (per-link mask, gradient window,
// pair list) ever goes to L2 / HBM and the whole stage is a single launch:
//   A  windows   35x35 window of every link's depth plane that reaches into the tile -> triangle ids in shared memory
//                (all 256 threads load, every load of the phase in flight at once), row coverage masks by ballot
//   B  pairs     silhouette pixel pairs by XOR of neighbouring coverage masks (one warp per link, lane = row) -> ONE pair
//                list for the tile
//   C  weights   blend weight of every pair (ehb_aa_pair), all threads stride over the tile's list: the work of a tile is
//                balanced over its 8 warps whatever its links look like
//   D  masks     per link: coverage as floats + the four contribution kinds in the reference's order
//   E  compose   S = min(sum of the link masks in link order, 1) -> staging tile -> ONE TMA tensor store (UTMASTG);
//                (S - ref)^2 -> loss; g = dL/dsum of the out region stays in shared memory
//   F  backward  every owned pair with a non-zero weight: analytic gradient of its two edge vertices
//                (ehb_aa_pair_grad), contracted with [x y z 1] on the fly (fp64), reduced per link by shuffles and
//                shared-memory accumulators, 12 fp64 atomics per (tile, link) into d loss / d mvp
// Links are processed in rounds of at most EHB_RL (their windows' storage), a round's pairs must fit the pair arrays
// (capacity = a launch parameter): a tile that needs several rounds runs A-D per round for the forward, and A-C again
// per round for the backward (g is only known once every link has been composed).  Tiles of a robot arm need one round.
#pragma once
#include "ehb_kernels.cuh"

#define EHB_MROWS (EHB_T + 1)            // out region of a tile: interior + one row / column on the high side
#define EHB_MW 36                        // row pitch (floats) of a link mask / the S / g window
#define EHB_MSZ (EHB_MROWS * EHB_MW)
#ifndef EHB_RL
#define EHB_RL 5                         // links whose windows are resident at a time
#endif
#define EHB_TTHREADS 256
#define EHB_TWARPS (EHB_TTHREADS / 32)
#define EHB_IDS_WORDS (EHB_NP + 3)       // ids of one window (35 x 35); the same storage later holds the link's mask
static_assert(EHB_IDS_WORDS >= EHB_MSZ, "the mask of a link reuses its window's storage");
static_assert(EHB_RL <= EHB_TWARPS, "one warp per resident link in the per-link phases");

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

struct EhbSlot {                         // one resident link of the tile
    int link;
    int x0, y0, w, h;                    // its depth plane
    long long off;
    int pairBase, nPairs;                // its part of the tile's pair list
};

// dynamic shared memory of a CTA (byte offsets; every block 16-byte aligned, the staging tile 128-byte aligned)
struct EhbTileSmem {
    float* stage;                        // [32][32]   composed tile, source of the TMA store
    float* S;                            // [33][36]   running sum of the link masks, then g = dL/dsum
    uint32_t* ids;                       // [RL][IDS_WORDS]
    unsigned long long* cov;             // [RL][36]
    float* alpha;                        // [cap]
    uint32_t* ptri;                      // [cap]
    unsigned short* pk;                  // [cap]  idx (11) | d << 11 | own << 12 | side << 13 | di << 14
    unsigned char* pslot;                // [cap]
    double* gacc;                        // [RL][12]
    EhbSlot* slot;                       // [RL]
    double* lsum;                        // [TWARPS]
    int* misc;                           // [8]
};
__host__ __device__ inline size_t ehb_tile_smem_bytes(int cap)
{
    size_t n = 4096 + EHB_MSZ * 4 + (size_t)EHB_RL * EHB_IDS_WORDS * 4 + (size_t)EHB_RL * 36 * 8;
    n = (n + 15) & ~(size_t)15;
    n += (size_t)cap * (4 + 4 + 2 + 1);
    n = (n + 15) & ~(size_t)15;
    n += EHB_RL * 12 * 8 + EHB_RL * sizeof(EhbSlot) + EHB_TWARPS * 8 + 8 * 4 + 64;
    return n + 128;                      // slack for the 128-byte alignment of the base
}
__device__ __forceinline__ EhbTileSmem ehb_tile_smem(unsigned char* base, int cap)
{
    base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(base) + 127) & ~(uintptr_t)127);
    EhbTileSmem s;
    s.stage = reinterpret_cast<float*>(base); base += 4096;
    s.S = reinterpret_cast<float*>(base); base += EHB_MSZ * 4;
    s.ids = reinterpret_cast<uint32_t*>(base); base += (size_t)EHB_RL * EHB_IDS_WORDS * 4;
    s.cov = reinterpret_cast<unsigned long long*>((reinterpret_cast<uintptr_t>(base) + 7) & ~(uintptr_t)7);
    base = reinterpret_cast<unsigned char*>(s.cov) + (size_t)EHB_RL * 36 * 8;
    base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(base) + 15) & ~(uintptr_t)15);
    s.alpha = reinterpret_cast<float*>(base); base += (size_t)cap * 4;
    s.ptri = reinterpret_cast<uint32_t*>(base); base += (size_t)cap * 4;
    s.pk = reinterpret_cast<unsigned short*>(base); base += (size_t)cap * 2;
    s.pslot = base; base += (size_t)cap;
    base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(base) + 15) & ~(uintptr_t)15);
    s.gacc = reinterpret_cast<double*>(base); base += EHB_RL * 12 * 8;
    s.slot = reinterpret_cast<EhbSlot*>(base); base += EHB_RL * sizeof(EhbSlot);
    base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(base) + 7) & ~(uintptr_t)7);
    s.lsum = reinterpret_cast<double*>(base); base += EHB_TWARPS * 8;
    s.misc = reinterpret_cast<int*>(base);
    return s;
}

// link index of the k-th set bit of `bits`
__device__ __forceinline__ int ehb_nth_bit(uint32_t bits, int k) { return (int)__fns(bits, 0, k + 1); }

// Everything a CTA knows about the tile it is working on.
struct EhbTileCtx {
    EhbTileSmem sm;
    int cap, tid, lane, warp;
    int item, tile, x0, y0, rx0, ry0, nl;
    uint32_t bits;
    bool needAA;
    int ow;
};

// A + B + C for the links [lNext, lNext + take) of the tile (in link order).  On return `take` is the number of links
// that fit the pair arrays (>= 1) and the pair list holds their `nPairsRound` pairs with weights.
__device__ __forceinline__ void ehb_tile_round(const EhbRobot& rb, const EhbParams& p, const EhbTileCtx& c, int lNext, int& take,
                                               int& nPairsRound)
{
    const EhbTileSmem& sm = c.sm;
    const int tid = c.tid, lane = c.lane, warp = c.warp, cap = c.cap;
    const int H = p.H, W = p.W, hlo = 1, ow = c.ow;
    __syncthreads();                                     // the previous round / tile is done with the arrays
    take = min(EHB_RL, c.nl - lNext);
    // ================================ A: slots, windows, coverage ================================
    if (tid < take) {
        const int l = ehb_nth_bit(c.bits, lNext + tid);
        const EhbPlane pl = p.plane[(size_t)c.item * p.L + l];
        EhbSlot s;
        s.link = l; s.x0 = pl.x0; s.y0 = pl.y0; s.w = pl.w; s.h = pl.h; s.off = pl.off; s.pairBase = 0; s.nPairs = 0;
        sm.slot[tid] = s;
    }
    __syncthreads();
    {
        const int total = take * EHB_NP;
        for (int e0 = tid; e0 < total; e0 += 6 * EHB_TTHREADS) {
            unsigned long long v[6];
#pragma unroll
            for (int k = 0; k < 6; k++) {       // every load of the batch is issued before the first is consumed
                v[k] = EHB_EMPTY;
                const int ee = e0 + k * EHB_TTHREADS;
                if (ee < total) {
                    const int s = ee / EHB_NP, i = ee - s * EHB_NP;
                    const int r = i / EHB_RS, cc = i - r * EHB_RS;
                    const EhbSlot& sl = sm.slot[s];
                    const int cx = c.rx0 + cc - sl.x0, cy = c.ry0 + r - sl.y0;
                    if ((unsigned)cx < (unsigned)sl.w && (unsigned)cy < (unsigned)sl.h)
                        v[k] = p.pool[sl.off + (long long)cy * sl.w + cx];
                }
            }
#pragma unroll
            for (int k = 0; k < 6; k++) {
                const int ee = e0 + k * EHB_TTHREADS;
                if (ee < total) {
                    const int s = ee / EHB_NP, i = ee - s * EHB_NP;
                    sm.ids[s * EHB_IDS_WORDS + i] = (uint32_t)v[k];   // low word = triangle id (all ones: empty)
                }
            }
        }
    }
    __syncthreads();
    // row coverage masks: (slot, row) pairs over the warps; bits 0..31 from one ballot, 32..34 from a second
    for (int sr = warp; sr < take * (EHB_RS + 1); sr += EHB_TWARPS) {
        const int s = sr / (EHB_RS + 1), r = sr - s * (EHB_RS + 1);
        unsigned long long m = 0ull;
        if (r < EHB_RS) {
            const uint32_t* row = sm.ids + s * EHB_IDS_WORDS + r * EHB_RS;
            const unsigned b0 = __ballot_sync(0xffffffffu, row[lane] != 0xFFFFFFFFu);
            const unsigned b1 = __ballot_sync(0xffffffffu, lane < EHB_RS - 32 && row[32 + lane] != 0xFFFFFFFFu);
            m = (unsigned long long)b0 | ((unsigned long long)b1 << 32);
        }
        if (lane == 0) sm.cov[s * 36 + r] = m;
    }
    __syncthreads();
    // ================================ B: silhouette pairs, one warp per slot, lane = window row ================================
    // columns whose pixel is inside the image, and for which the right neighbour is too
    const unsigned long long inX = ehb_bits(-c.rx0, W - 1 - c.rx0), inX1 = ehb_bits(-c.rx0, W - 2 - c.rx0);
    unsigned long long hm[2] = {0ull, 0ull}, vm[2] = {0ull, 0ull}, om[2] = {0ull, 0ull};
    int o0 = 0, o1 = 0;
    if (warp < take) {
        const unsigned long long* cv = sm.cov + warp * 36;
        int cnt[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int r = lane + 32 * h;
            if (r < EHB_RS) {
                const int py = c.ry0 + r;
                const unsigned long long cm = cv[r], cu = cv[r + 1];
                // pairs wanted: forward = those touching a pixel of the out region; otherwise only owned ones
                unsigned long long wantH, wantV;
                if (c.needAA) {
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
        o0 = inc - cnt[0]; o1 = tot0 + inc1 - cnt[1];
        if (lane == 0) sm.slot[warp].nPairs = tot0 + tot1;
    }
    __syncthreads();
    // fit the round into the pair arrays: the longest prefix of slots whose pairs fit (at least one slot)
    nPairsRound = 0;
    int fit = 0;
    for (int s = 0; s < take; s++) {
        const int n = sm.slot[s].nPairs;
        if (s > 0 && nPairsRound + n > cap) break;
        nPairsRound += n; fit = s + 1;
    }
    if (nPairsRound > cap) {          // one window with more pairs than the arrays hold: flagged, grow and rerun
        if (tid == 0) atomicOr(&p.ctr->flags, 1u);
        nPairsRound = cap;
    }
    take = fit;
    __syncthreads();
    if (tid == 0) {
        int b = 0;
        for (int s = 0; s < take; s++) { sm.slot[s].pairBase = b; b += sm.slot[s].nPairs; }
    }
    __syncthreads();
    if (warp < take) {
        const int base = sm.slot[warp].pairBase;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            int o = base + (h ? o1 : o0);
            unsigned long long hxm = hm[h], vym = vm[h];
            const uint32_t rowBase = (uint32_t)(lane + 32 * h) * EHB_RS;
            while (hxm) {
                const int b = __ffsll((long long)hxm) - 1;
                hxm &= hxm - 1;
                if (o < cap) {
                    sm.pk[o] = (unsigned short)((rowBase + b) | (((om[h] >> b) & 1ull) ? (1u << 12) : 0u));
                    sm.pslot[o] = (unsigned char)warp;
                }
                o++;
            }
            while (vym) {
                const int b = __ffsll((long long)vym) - 1;
                vym &= vym - 1;
                if (o < cap) {
                    sm.pk[o] = (unsigned short)((rowBase + b) | (1u << 11) | (((om[h] >> b) & 1ull) ? (1u << 12) : 0u));
                    sm.pslot[o] = (unsigned char)warp;
                }
                o++;
            }
        }
    }
    __syncthreads();
    // ================================ C: blend weights, all threads over the tile's pair list ================================
    for (int i = tid; i < nPairsRound; i += EHB_TTHREADS) {
        const uint32_t pk = sm.pk[i];
        const int s = sm.pslot[i];
        const int l = sm.slot[s].link;
        const int idx = pk & 2047, d = (pk >> 11) & 1;
        const int ly = idx / EHB_RS, lx = idx - ly * EHB_RS;
        const uint32_t* idw = sm.ids + s * EHB_IDS_WORDS;
        const uint32_t ka = idw[idx], kb = idw[idx + (d ? EHB_RS : 1)];
        const int side = ka != 0xFFFFFFFFu ? 0 : 1;
        const uint32_t t = side ? kb : ka;
        int di;
        const float al = ehb_aa_pair(rb.link[l], p.vclip + (size_t)c.item * p.Vtot + rb.voff[l], (int)t, side, c.rx0 + lx, c.ry0 + ly, d,
                                     H, W, &di);
        sm.alpha[i] = al;
        sm.ptri[i] = t;
        sm.pk[i] = (unsigned short)(pk | ((uint32_t)side << 13) | ((uint32_t)di << 14));
    }
    __syncthreads();
}

// D: the antialiased masks of the round's links (out region), then the running sum in link order.
__device__ __forceinline__ void ehb_tile_masks(const EhbTileCtx& c, int take, bool first)
{
    const EhbTileSmem& sm = c.sm;
    const int lane = c.lane, warp = c.warp, hlo = 1, ow = c.ow;
    // colour = coverage as floats (the window's ids are dead: their storage becomes the mask), then the pair
    // contributions.  A pixel receives at most one contribution of each kind and the reference adds them in the order
    // pair(p,p+x), pair(p,p+y), pair(p-x,p), pair(p-y,p): four sweeps over the link's pairs, one kind each (receiver = p0
    // when alpha > 0, p1 otherwise; the contribution is alpha * (colour[p1] - colour[p0])).
    if (warp < take) {
        float* ot = reinterpret_cast<float*>(sm.ids + warp * EHB_IDS_WORDS);
        const unsigned long long* cv = sm.cov + warp * 36;
        for (int i = lane; i < ow * EHB_MW; i += 32) {
            const int qy = i / EHB_MW, qx = i - qy * EHB_MW;
            ot[i] = (qx < ow && ((cv[hlo + qy] >> (hlo + qx)) & 1ull)) ? 1.f : 0.f;
        }
        const int pb = sm.slot[warp].pairBase, pn = max(0, min(sm.slot[warp].nPairs, c.cap - pb));
#pragma unroll 1
        for (int kind = 0; kind < 4; kind++) {
            __syncwarp();
            for (int i = pb + lane; i < pb + pn; i += 32) {
                const uint32_t pk = sm.pk[i];
                const float al = sm.alpha[i];
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
    }
    __syncthreads();
    // running sum in link order (rb_solver.py:68): S = m_first, then S = S + m_l
    for (int i = c.tid; i < ow * EHB_MW; i += EHB_TTHREADS) {
        float s = first ? 0.f : sm.S[i];
        for (int k = 0; k < take; k++) {
            const float m = reinterpret_cast<const float*>(sm.ids + k * EHB_IDS_WORDS)[i];
            s = (first && k == 0) ? m : s + m;
        }
        sm.S[i] = s;
    }
}

// F: backward of the resident pairs (sm.S holds g = dL/dsum of the out region).
__device__ __forceinline__ void ehb_tile_backward(const EhbRobot& rb, const EhbParams& p, const EhbTileCtx& c, int take, int nPairsRound)
{
    const EhbTileSmem& sm = c.sm;
    const int tid = c.tid, lane = c.lane, hlo = 1;
    if (tid < EHB_RL * 12) sm.gacc[tid] = 0.0;
    __syncthreads();                                     // ... and g / the pair weights are complete
    for (int i0 = 0; i0 < nPairsRound; i0 += EHB_TTHREADS) {
        const int i = i0 + tid;
        double acc[12];
#pragma unroll
        for (int k = 0; k < 12; k++) acc[k] = 0.0;
        int key = -1;
        if (i < nPairsRound) {
            const uint32_t pk = sm.pk[i];
            const float al = sm.alpha[i];
            if ((pk & (1u << 12)) && al != 0.f) {        // owned (p0 inside the tile's interior) with a non-zero weight
                const int idx = pk & 2047, d = (pk >> 11) & 1, side = (pk >> 13) & 1, di = (pk >> 14) & 3;
                const int ridx = al > 0.f ? idx : idx + (d ? EHB_RS : 1);
                const int ry = ridx / EHB_RS, rxw = ridx - ry * EHB_RS;
                const float g = sm.S[(ry - hlo) * EHB_MW + (rxw - hlo)];
                const float dd = g * (side ? 1.f : -1.f);   // g * (c1 - c0)
                if (dd != 0.f) {
                    const int s = sm.pslot[i];
                    const int l = sm.slot[s].link;
                    const int ly = idx / EHB_RS, lx = idx - ly * EHB_RS;
                    const EhbLink& lk = rb.link[l];
                    int vi1, vi2;
                    float g1[3], g2[3];
                    ehb_aa_pair_grad(lk, p.vclip + (size_t)c.item * p.Vtot + rb.voff[l], (int)sm.ptri[i], side, di, al, dd, c.rx0 + lx,
                                     c.ry0 + ly, d, p.H, p.W, &vi1, &vi2, g1, g2);
                    const float4 va = __ldg(lk.verts + vi1), vb = __ldg(lk.verts + vi2);
                    const double ha[4] = {(double)va.x, (double)va.y, (double)va.z, 1.0};
                    const double hb[4] = {(double)vb.x, (double)vb.y, (double)vb.z, 1.0};
#pragma unroll
                    for (int rr = 0; rr < 3; rr++)
#pragma unroll
                        for (int cc = 0; cc < 4; cc++) acc[4 * rr + cc] = (double)g1[rr] * ha[cc] + (double)g2[rr] * hb[cc];
                    key = s;
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
        // everything of a warp that belongs to one link: reduced by shuffles, then one shared-memory atomic per component
        unsigned todo = __ballot_sync(0xffffffffu, key >= 0);
        while (todo) {
            const int leader = __ffs(todo) - 1;
            const int k0 = __shfl_sync(0xffffffffu, key, leader);
            const bool mine = key == k0;
            todo &= ~__ballot_sync(0xffffffffu, mine);
#pragma unroll
            for (int k = 0; k < 12; k++) {
                const double v = ehb_warp_sum(mine ? acc[k] : 0.0);
                if (lane == 0 && v != 0.0) atomicAdd(sm.gacc + k0 * 12 + k, v);
            }
        }
    }
    __syncthreads();
    if (tid < take * 12 && p.gmvp) {
        const int s = tid / 12, k = tid - s * 12;
        const double v = sm.gacc[tid];
        // rows x (0), y (1), w (3) of d loss / d mvp; the z row carries no gradient
        if (v != 0.0) atomicAdd(p.gmvp + ((size_t)c.item * p.L + sm.slot[s].link) * 16 + (k < 8 ? k : k + 4), v);
    }
}

__global__ void __launch_bounds__(EHB_TTHREADS) ehb_k_tiles(const __grid_constant__ EhbRobot rb,
                                                            const __grid_constant__ EhbParams p, int cap)
{
    ehb_pdl_enter();
    extern __shared__ unsigned char ehb_dsm[];
    EhbTileCtx c;
    c.sm = ehb_tile_smem(ehb_dsm, cap);
    c.cap = cap;
    c.tid = threadIdx.x; c.lane = c.tid & 31; c.warp = c.tid >> 5;
    const EhbTileSmem& sm = c.sm;
    const int tid = c.tid, lane = c.lane, warp = c.warp;
    const int H = p.H, W = p.W;
    const bool fused = p.mode == EHB_MODE_FUSED;
    c.needAA = fused || p.mode == EHB_MODE_AA_FWD;
    const bool doBwd = (fused && p.do_bwd) || p.mode == EHB_MODE_AA_BWD;
    const int oext = (fused && p.do_bwd) ? 1 : 0;
    c.ow = EHB_T + oext;                                 // out region whose S is needed
    const int ow = c.ow;
    const int refKind = !fused ? 0 : (p.refBits ? 3 : (p.ref ? 1 : (p.ref_u8 ? 2 : 0)));
    const unsigned nHeavy = p.ctr->nTiles, nEntries = nHeavy + p.ctr->nLight;
    const uint32_t linkMask = p.L >= 32 ? 0xFFFFFFFFu : ((1u << p.L) - 1u);
    const bool tma = p.masks != nullptr && p.useTma;
    bool storePending = false;                           // thread 0: a TMA store may still be reading the staging tile

    for (unsigned e = blockIdx.x; e < nEntries; e += gridDim.x) {
        const uint32_t wid = ehb_list_at(p, e, nHeavy);
        c.item = (int)(wid / (uint32_t)p.ntiles); c.tile = (int)(wid - (uint32_t)c.item * (uint32_t)p.ntiles);
        c.x0 = (c.tile % p.ntx) * EHB_T; c.y0 = (c.tile / p.ntx) * EHB_T;
        c.rx0 = c.x0 - 1; c.ry0 = c.y0 - 1;              // window = tile + 1 low / 2 high halo pixels (35 x 35), all modes
        c.bits = p.touch[wid] & linkMask;
        c.nl = __popc(c.bits);
        const int item = c.item, x0 = c.x0, y0 = c.y0, nl = c.nl;
        const size_t ibase = (size_t)item * H * W;
        // ---- reference values of this thread's out-region pixels: issued now, consumed in E -------------------------
        // pixel k of a thread: i = tid + k * 256 over the 33 x 33 out region (row-major), 5 per thread at most
        float rf[5];
#pragma unroll
        for (int k = 0; k < 5; k++) {
            rf[k] = 0.f;
            const int i = tid + k * EHB_TTHREADS;
            if (refKind && i < EHB_MROWS * EHB_MROWS) {
                const int qy = i / EHB_MROWS, qx = i - qy * EHB_MROWS;
                const int px = x0 + qx, py = y0 + qy;
                if (px < W && py < H && qy < ow && qx < ow) {
                    if (refKind == 3) {
                        const uint32_t wbits = __ldg(p.refBits + ((size_t)item * H + py) * p.ntx + (px >> 5));
                        rf[k] = ((wbits >> (px & 31)) & 1u) ? 1.f : 0.f;
                    } else {
                        const size_t o = ibase + (size_t)(H - 1 - py) * W + px;
                        rf[k] = refKind == 1 ? __ldg(p.ref + o) : (__ldg(p.ref_u8 + o) ? 1.f : 0.f);
                    }
                }
            }
        }
        __syncthreads();                                 // the previous tile is done with S
        if (p.mode == EHB_MODE_AA_BWD) {                 // g = dL/dmask comes from the caller
            for (int i = tid; i < EHB_MROWS * EHB_MROWS; i += EHB_TTHREADS) {
                const int qy = i / EHB_MROWS, qx = i - qy * EHB_MROWS;
                const int px = x0 + qx, py = y0 + qy;
                sm.S[qy * EHB_MW + qx] = (px < W && py < H) ? __ldg(p.dy + ibase + (size_t)(H - 1 - py) * W + px) : 0.f;
            }
        }
        // ---- rounds over the tile's links (one round for the tiles of a robot arm) -------------------------------------
        int lNext = 0, rounds = 0, take = 0, nPairsRound = 0;
        bool first = true;
        while (lNext < nl) {
            ehb_tile_round(rb, p, c, lNext, take, nPairsRound);
            if (c.needAA) ehb_tile_masks(c, take, first);
            else ehb_tile_backward(rb, p, c, take, nPairsRound);          // operator backward: g is already there
            lNext += take; first = false; rounds++;
        }
        if (!c.needAA) continue;
        // ================================ E: compose, loss, dL/dsum ================================
        __syncthreads();
        if (tid == 0 && storePending) { ehb_bulk_wait_read(); storePending = false; }
        __syncthreads();
        double lacc = 0.0;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const int i = tid + k * EHB_TTHREADS;
            if (i >= EHB_MROWS * EHB_MROWS) continue;
            const int qy = i / EHB_MROWS, qx = i - qy * EHB_MROWS;
            if (qy >= ow || qx >= ow) continue;
            const int px = x0 + qx, py = y0 + qy;
            const bool inImg = px < W && py < H;
            const float s = nl > 0 ? sm.S[qy * EHB_MW + qx] : 0.f;
            const float Sv = (p.clamp && s > 1.f) ? 1.f : s;
            float gv = 0.f;
            if (refKind && inImg) {
                const float diff = Sv - rf[k];
                if (qx < EHB_T && qy < EHB_T) lacc += (double)(diff * diff);
                gv = (!p.clamp || s <= 1.f) ? (2.f * diff) * p.invB : 0.f;
            }
            if (qx < EHB_T && qy < EHB_T && p.masks) {
                if (tma) sm.stage[(EHB_T - 1 - qy) * EHB_T + qx] = Sv;          // image rows run downwards
                else if (inImg) p.masks[ibase + (size_t)(H - 1 - py) * W + px] = Sv;
            }
            if (oext) sm.S[qy * EHB_MW + qx] = gv;      // (a thread rewrites only the pixels it has just read)
        }
        if (refKind && p.loss) {
            lacc = ehb_warp_sum(lacc);
            if (lane == 0) sm.lsum[warp] = lacc;
        }
        if (tma) ehb_fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            if (tma) {
                // the tile's image rows: H - 32 - y0 .. H - 1 - y0 (negative start for the top row of tiles: clipped)
                ehb_tma_store_3d(&p.tmMask, sm.stage, x0, H - EHB_T - y0, item);
                ehb_bulk_commit();
                storePending = true;
            }
            if (refKind && p.loss) {
                double t = 0.0;
                for (int w = 0; w < EHB_TWARPS; w++) t += sm.lsum[w];
                if (refKind == 3) t -= (double)__ldg(p.refCnt + (size_t)item * p.ntiles + c.tile);   // loss[item] starts at sum(ref)
                if (t != 0.0) atomicAdd(&p.loss[item], t);
            }
        }
        if (!doBwd || nl == 0) continue;
        if (rounds == 1) {
            ehb_tile_backward(rb, p, c, take, nPairsRound);          // the pairs of the forward are still resident
        } else {
            lNext = 0;
            while (lNext < nl) {                                      // g is known now: rebuild each round's pairs for the backward
                ehb_tile_round(rb, p, c, lNext, take, nPairsRound);
                ehb_tile_backward(rb, p, c, take, nPairsRound);
                lNext += take;
            }
        }
    }
    if (tid == 0 && storePending) ehb_bulk_wait_read();
}
