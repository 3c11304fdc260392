// ehb_tiles_impl.cuh -- the body of the image-space kernel, included by ehb_tiles.cuh once per CTA size (EHB_TTHREADS,
// EHB_TMIN_BLOCKS, namespace EHB_TNS).  No include guard on purpose.
#undef EHB_TWARPS
#define EHB_TWARPS (EHB_TTHREADS / 32)
namespace EHB_TNS {

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
    int pairBase, nPairs;                // its part of the tile's pair list
    int pad;
};

struct __align__(128) EhbTileSm {
    float mbuf[EHB_TMB][EHB_MSZ];                // antialiased masks of EHB_TMB links at a time; [0] doubles as the 32 x 32 staging tile of the TMA store
    float S[EHB_MSZ];                            // running sum of the link masks, then g = dL/dsum
    unsigned long long cov[EHB_RL][36];          // row coverage masks of the round's windows (rows 0..34, [35] = 0)
    float alpha[EHB_CAPS];
    uint32_t ptri[EHB_CAPS];
    unsigned short pk[EHB_CAPS];                 // idx (11) | d << 11 | own << 12 | side << 13 | di << 14
    unsigned char pslot[EHB_CAPS];
    double gacc[EHB_RL][12];
    double lsum[EHB_TWARPS];
    EhbPlane planes[EHB_MAX_LINKS];              // the item's depth planes, fetched together with the tile's link bits
    EhbSlot slot[EHB_RL];
    uint32_t rbits[EHB_MROWS][2];                // registered reference: bits of the out region's rows
    uint32_t bits, refCnt;
    int slab;                                    // index of the CTA's slab in the pair pool (-1: none yet)
#ifdef EHB_STATS
    long long stat[8];                           // this tile's cycles per phase (A B C D E F whole) and its pairs
#endif
};

// the pair arrays of a round: shared memory, or a slab in global memory
struct EhbPairs {
    float* alpha; uint32_t* ptri; unsigned short* pk; unsigned char* pslot;
};

// Everything a CTA knows about the tile it is working on.
struct EhbTileCtx {
    int tid, lane, warp;
    int item, tile, x0, y0, rx0, ry0, nl;
    uint32_t bits;
};

// hm / vm / om of window row r of one slot: silhouette pairs to the right (hm) and upwards (vm) that are wanted, and the
// columns whose pairs this tile owns (om)
template <bool NEEDAA, int OW>
__device__ __forceinline__ void ehb_row_pairs(const unsigned long long* cv, int r, int py, int H, unsigned long long inX,
                                              unsigned long long inX1, unsigned long long& hm, unsigned long long& vm,
                                              unsigned long long& om)
{
    const int hlo = 1;
    hm = vm = om = 0ull;
    if (r >= EHB_RS) return;
    const unsigned long long cm = cv[r], cu = cv[r + 1];
    // pairs wanted: forward = those touching a pixel of the out region; otherwise only owned ones
    unsigned long long wantH, wantV;
    if (NEEDAA) {
        wantH = (r >= hlo && r <= hlo + OW - 1) ? ehb_bits(hlo - 1, hlo + OW - 1) : 0ull;
        wantV = (r >= hlo - 1 && r <= hlo + OW - 1) ? ehb_bits(hlo, hlo + OW - 1) : 0ull;
    } else {
        wantH = wantV = (r >= hlo && r <= hlo + EHB_T - 1) ? ehb_bits(hlo, hlo + EHB_T - 1) : 0ull;
    }
    const bool rowIn = py >= 0 && py < H;
    if (rowIn) hm = (cm ^ (cm >> 1)) & inX1 & wantH & ehb_bits(0, EHB_RS - 2);
    if (rowIn && py < H - 1 && r < EHB_RS - 1) vm = (cm ^ cu) & inX & wantV;
    om = (r >= hlo && r <= hlo + EHB_T - 1) ? ehb_bits(hlo, hlo + EHB_T - 1) : 0ull;
}

// A + B + C for the links [lNext, lNext + take) of the tile (in link order).  On return `take` is the number of links
// whose pairs fit one list (>= 1), `pr` the list (shared memory or a global slab) with the `nPairsRound` pairs and weights.
template <bool NEEDAA, int OW>
__device__ __forceinline__ void ehb_tile_round(const EhbRobot& rb, const EhbParams& p, EhbTileSm& sm, const EhbTileCtx& c, int lNext,
                                               int& take, int& nPairsRound, EhbPairs& pr)
{
    const int tid = c.tid, lane = c.lane, warp = c.warp;
    const int H = p.H, W = p.W;
    __syncthreads();                                     // the previous round / tile is done with the arrays
    EHB_STAT_T(tA);
    take = min(EHB_RL, c.nl - lNext);
    // ================================ A: coverage of the windows ================================
    // The rasterizer kept one coverage bit per pixel of every plane (ehb_bits_set): a window row is 35 bits out of two
    // 64-bit words.  A warp per link, lane = window row (rows 32..34 by the first three lanes): every load of the round is
    // in flight at once, and a link costs 0.6 KB of L2 traffic instead of the 9.8 KB of its depth-plane window (with every
    // tile of the pass starting together, those windows queued at L2 for 5,000 cycles per link).
    {
        uint32_t b = c.bits;
        for (int q = 0; q < lNext; q++) b &= b - 1;
        for (int s = 0; s < take; s++) {
            const int l = __ffs(b) - 1;
            b &= b - 1;
            if (tid == 0) { sm.slot[s].link = l; sm.slot[s].nPairs = 0; sm.slot[s].pairBase = 0; }
            if ((s & (EHB_TWARPS - 1)) != warp) continue;
            const EhbPlane& pl = sm.planes[l];
            const int bw = (pl.w + 63) >> 6;
            const int bx0 = c.rx0 - pl.x0;                    // plane column of window column 0 (may be negative)
            const int s0 = max(bx0, 0), wi = s0 >> 6, bo = s0 & 63;
#pragma unroll
            for (int part = 0; part < 2; part++) {
                const int r = part ? 32 + lane : lane;
                if (r > EHB_RS) break;
                unsigned long long v = 0ull;
                const int by = c.ry0 + r - pl.y0;
                if (r < EHB_RS && (unsigned)by < (unsigned)pl.h && wi < bw && bx0 < pl.w && bx0 > -EHB_RS) {
                    const unsigned long long* row = p.bits + pl.boff + (long long)by * bw;
                    const unsigned long long lo = __ldcg(row + wi), hi = (wi + 1 < bw && bo) ? __ldcg(row + wi + 1) : 0ull;
                    v = bo ? ((lo >> bo) | (hi << (64 - bo))) : lo;
                    if (bx0 < 0) v <<= -bx0;                  // window columns left of the plane are empty
                    v &= (1ull << EHB_RS) - 1ull;
                }
                if (r <= EHB_RS) sm.cov[s][r] = v;            // (row 35 = 0: the pair search looks one row up)
            }
        }
    }
    __syncthreads();
    EHB_STAT_T(tB);
    EHB_STAT_ADD(5, tB - tA);
    // ================================ B: silhouette pairs, a warp per slot, lane = window row ================================
    // columns whose pixel is inside the image, and for which the right neighbour is too
    const unsigned long long inX = ehb_bits(-c.rx0, W - 1 - c.rx0), inX1 = ehb_bits(-c.rx0, W - 2 - c.rx0);
    for (int s = warp; s < take; s += EHB_TWARPS) {       // count
        int cnt = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            unsigned long long hm, vm, om;
            ehb_row_pairs<NEEDAA, OW>(sm.cov[s], lane + 32 * h, c.ry0 + lane + 32 * h, H, inX, inX1, hm, vm, om);
            cnt += __popcll(hm) + __popcll(vm);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) sm.slot[s].nPairs = cnt;
    }
    __syncthreads();
    // fit the round into one pair list: the longest prefix of slots whose pairs fit a slab (at least one slot: a window
    // has at most 2380 pairs); every thread computes the same answer from the slots' counts
    nPairsRound = 0;
    int fit = 0;
    for (int s = 0; s < take; s++) {
        const int n = sm.slot[s].nPairs;
        if (s > 0 && nPairsRound + n > EHB_CAPG) break;
        nPairsRound += n; fit = s + 1;
    }
    nPairsRound = min(nPairsRound, EHB_CAPG);
    take = fit;
    pr.alpha = sm.alpha; pr.ptri = sm.ptri; pr.pk = sm.pk; pr.pslot = sm.pslot;
    if (nPairsRound > EHB_CAPS) {
        // more pairs than shared memory holds: this CTA takes a slab of the context's pair pool (kept for its later rounds)
        if (tid == 0 && sm.slab < 0) {
            const unsigned k = atomicAdd(&p.ctr->slabCursor, 1u);
            if (k < (unsigned)p.nSlabs) sm.slab = (int)k;
            else ehb_raise(p, 1u);              // pool exhausted: flagged, the host grows it and reruns the pass
        }
        __syncthreads();
        EHB_STAT_ADD(12, 1);
        if (sm.slab >= 0) {
            unsigned char* base = p.pairPool + (size_t)sm.slab * EHB_SLAB_BYTES;
            pr.alpha = reinterpret_cast<float*>(base);
            pr.ptri = reinterpret_cast<uint32_t*>(base + EHB_CAPG * 4);
            pr.pk = reinterpret_cast<unsigned short*>(base + EHB_CAPG * 8);
            pr.pslot = base + EHB_CAPG * 10;
        } else {
            nPairsRound = min(nPairsRound, EHB_CAPS);      // (results of this pass are discarded)
        }
    }
    const int capNow = pr.alpha == sm.alpha ? EHB_CAPS : EHB_CAPG;
    for (int s = warp; s < take; s += EHB_TWARPS) {       // write: the rows' pairs in row order, horizontal before vertical
        int base = 0;
        for (int q = 0; q < s; q++) base += sm.slot[q].nPairs;
        if (lane == 0) sm.slot[s].pairBase = base;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            unsigned long long hm, vm, om;
            ehb_row_pairs<NEEDAA, OW>(sm.cov[s], lane + 32 * h, c.ry0 + lane + 32 * h, H, inX, inX1, hm, vm, om);
            const int n = __popcll(hm) + __popcll(vm);
            int inc = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            int o = base + inc - n;
            base += __shfl_sync(0xffffffffu, inc, 31);
            const uint32_t rowBase = (uint32_t)(lane + 32 * h) * EHB_RS;
            while (hm) {
                const int bb = __ffsll((long long)hm) - 1;
                hm &= hm - 1;
                if (o < capNow) {
                    pr.pk[o] = (unsigned short)((rowBase + bb) | (((om >> bb) & 1ull) ? (1u << 12) : 0u));
                    pr.pslot[o] = (unsigned char)s;
                }
                o++;
            }
            while (vm) {
                const int bb = __ffsll((long long)vm) - 1;
                vm &= vm - 1;
                if (o < capNow) {
                    pr.pk[o] = (unsigned short)((rowBase + bb) | (1u << 11) | (((om >> bb) & 1ull) ? (1u << 12) : 0u));
                    pr.pslot[o] = (unsigned char)s;
                }
                o++;
            }
        }
    }
    __syncthreads();
    EHB_STAT_T(tC);
    EHB_STAT_ADD(6, tC - tB);
    EHB_STAT_ADD(3, nPairsRound);
    // ================================ C: blend weights, all threads over the tile's pair list ================================
    for (int i = tid; i < nPairsRound; i += EHB_TTHREADS) {
        const uint32_t pk = pr.pk[i];
        const int s = pr.pslot[i];
        const int l = sm.slot[s].link;
        const int idx = pk & 2047, d = (pk >> 11) & 1;
        const int ly = idx / EHB_RS, lx = idx - ly * EHB_RS;
        // the covered pixel of the pair shows the triangle: p0 when it is covered, else p1 (the plane was read by this SM
        // in A: an L1 / L2 hit)
        const int side = ((sm.cov[s][ly] >> lx) & 1ull) ? 0 : 1;
        const int wx = lx + (side ? 1 - d : 0), wy = ly + (side ? d : 0);
        const EhbPlane& pl = sm.planes[l];
        const uint32_t t = (uint32_t)p.pool[pl.off + (long long)(c.ry0 + wy - pl.y0) * pl.w + (c.rx0 + wx - pl.x0)];
        int di;
        const float al = ehb_aa_pair(rb.link[l], p.vclip + (size_t)c.item * p.Vtot + rb.voff[l], (int)t, side, c.rx0 + lx, c.ry0 + ly, d,
                                     H, W, &di);
        pr.alpha[i] = al;
        pr.ptri[i] = t;
        pr.pk[i] = (unsigned short)(pk | ((uint32_t)side << 13) | ((uint32_t)di << 14));
    }
    __syncthreads();
    EHB_STAT_ADD(7, clock64() - tC);
}

// D: the antialiased masks of the round's links, EHB_TMB links at a time in as many mask buffers, then the running sum in
// link order (rb_solver.py:68): S = m_first, S = S + m_l.  A pixel receives at most one contribution of each kind and the
// reference adds them in the order pair(p,p+x), pair(p,p+y), pair(p-x,p), pair(p-y,p): four sweeps, one kind each, over the
// pairs of ALL links of the group (receiver = p0 when alpha > 0, p1 otherwise; contribution alpha * (colour[p1] - colour[p0]))
// -- the seven block barriers of the phase are paid once per group instead of once per link (2,600 cycles per link).
template <int OW>
__device__ __forceinline__ void ehb_tile_masks(const EhbParams& p, EhbTileSm& sm, const EhbTileCtx& c, const EhbPairs& pr, int take,
                                               int nPairsRound, bool first)
{
    EHB_STAT_T(tD);
    const int tid = c.tid, lane = c.lane, warp = c.warp, hlo = 1;
    for (int s0 = 0; s0 < take; s0 += EHB_TMB, first = false) {
        const int ns = min(EHB_TMB, take - s0);
        // colour = coverage as floats: the warps take (link, row)s, lane = column (+ columns 32 .. 35 by the first lanes)
        for (int k = warp; k < ns * OW; k += EHB_TWARPS) {
            const int sl = k / OW, qy = k - sl * OW;
            const unsigned long long cw = sm.cov[s0 + sl][hlo + qy] >> hlo;
            float* row = sm.mbuf[sl] + qy * EHB_MW;
            row[lane] = ((cw >> lane) & 1ull) ? 1.f : 0.f;
            if (lane < EHB_MW - 32) row[32 + lane] = (32 + lane < OW && (((cw >> 32) >> lane) & 1ull)) ? 1.f : 0.f;
        }
        const int lo = sm.slot[s0].pairBase;
        const int hi = min(sm.slot[s0 + ns - 1].pairBase + sm.slot[s0 + ns - 1].nPairs, nPairsRound);
        if (hi > lo) {
#pragma unroll 1
            for (int kind = 0; kind < 4; kind++) {
                __syncthreads();
                for (int i = lo + tid; i < hi; i += EHB_TTHREADS) {
                    const uint32_t pk = pr.pk[i];
                    const float al = pr.alpha[i];
                    const int d = (pk >> 11) & 1;
                    const bool pos = al > 0.f;
                    if (al == 0.f || d != (kind & 1) || pos != (kind < 2)) continue;
                    const int idx = pk & 2047, side = (pk >> 13) & 1;
                    const int ridx = pos ? idx : idx + (d ? EHB_RS : 1);
                    const int ry = ridx / EHB_RS, rxw = ridx - ry * EHB_RS;
                    const int qy = ry - hlo, qx = rxw - hlo;
                    if (qy < 0 || qx < 0 || qy >= OW || qx >= OW) continue;
                    const float delta = side ? 1.f : -1.f;   // colour[p1] - colour[p0]: p1 is the covered one when side = 1
                    sm.mbuf[(int)pr.pslot[i] - s0][qy * EHB_MW + qx] += al * delta;
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < OW * EHB_MW; i += EHB_TTHREADS) {
            float acc = first ? sm.mbuf[0][i] : sm.S[i] + sm.mbuf[0][i];
            for (int sl = 1; sl < ns; sl++) acc = acc + sm.mbuf[sl][i];
            sm.S[i] = acc;
        }
        __syncthreads();
    }
    EHB_STAT_ADD(8, clock64() - tD);
}

// F: backward of the resident pairs (sm.S holds g = dL/dsum of the out region).
__device__ __forceinline__ void ehb_tile_backward(const EhbRobot& rb, const EhbParams& p, EhbTileSm& sm, const EhbTileCtx& c,
                                                  const EhbPairs& pr, int take, int nPairsRound)
{
    EHB_STAT_T(tF);
    const int tid = c.tid, lane = c.lane, hlo = 1;
    if (tid < EHB_RL * 12) (&sm.gacc[0][0])[tid] = 0.0;
    __syncthreads();                                     // ... and g / the pair weights are complete
    for (int i0 = 0; i0 < nPairsRound; i0 += EHB_TTHREADS) {
        const int i = i0 + tid;
        double acc[12];
#pragma unroll
        for (int k = 0; k < 12; k++) acc[k] = 0.0;
        int key = -1;
        if (i < nPairsRound) {
            const uint32_t pk = pr.pk[i];
            const float al = pr.alpha[i];
            if ((pk & (1u << 12)) && al != 0.f) {        // owned (p0 inside the tile's interior) with a non-zero weight
                const int idx = pk & 2047, d = (pk >> 11) & 1, side = (pk >> 13) & 1, di = (pk >> 14) & 3;
                const int ridx = al > 0.f ? idx : idx + (d ? EHB_RS : 1);
                const int ry = ridx / EHB_RS, rxw = ridx - ry * EHB_RS;
                const float g = sm.S[(ry - hlo) * EHB_MW + (rxw - hlo)];
                const float dd = g * (side ? 1.f : -1.f);   // g * (c1 - c0)
                if (dd != 0.f) {
                    const int s = pr.pslot[i];
                    const int l = sm.slot[s].link;
                    const int ly = idx / EHB_RS, lx = idx - ly * EHB_RS;
                    const EhbLink& lk = rb.link[l];
                    int vi1, vi2;
                    float g1[3], g2[3];
                    ehb_aa_pair_grad(lk, p.vclip + (size_t)c.item * p.Vtot + rb.voff[l], (int)pr.ptri[i], side, di, al, dd, c.rx0 + lx,
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
        // everything of a warp that belongs to one link: the 12 components are reduced TOGETHER (a reduce-scatter: every
        // step halves the values a lane holds and doubles the lanes they are summed over -- 16 exchanges instead of the 60 of
        // twelve butterflies), component c ends up on lane 2c, and those twelve lanes issue their shared-memory atomics at
        // once (twelve one after the other from lane 0, each a compare-and-swap loop, were half of this phase)
        unsigned todo = __ballot_sync(0xffffffffu, key >= 0);
        while (todo) {
            const int leader = __ffs(todo) - 1;
            const int k0 = __shfl_sync(0xffffffffu, key, leader);
            const bool mine = key == k0;
            todo &= ~__ballot_sync(0xffffffffu, mine);
            double v8[8], v4[4], v2[2], v1;
            const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
            for (int j = 0; j < 8; j++) {              // component index = 8 * b4 + j
                const double lo = mine ? acc[j] : 0.0, hi = (mine && j + 8 < 12) ? acc[(j + 8) % 12] : 0.0;
                const double keep = b4 ? hi : lo, give = b4 ? lo : hi;
                v8[j] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {              // ... + 4 * b3 + j
                const double keep = b3 ? v8[j + 4] : v8[j], give = b3 ? v8[j] : v8[j + 4];
                v4[j] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
            }
#pragma unroll
            for (int j = 0; j < 2; j++) {              // ... + 2 * b2 + j
                const double keep = b2 ? v4[j + 2] : v4[j], give = b2 ? v4[j] : v4[j + 2];
                v2[j] = keep + __shfl_xor_sync(0xffffffffu, give, 4);
            }
            {                                           // ... + b1
                const double keep = b1 ? v2[1] : v2[0], give = b1 ? v2[0] : v2[1];
                v1 = keep + __shfl_xor_sync(0xffffffffu, give, 2);
            }
            v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
            const int comp = lane >> 1;                // = 8 b4 + 4 b3 + 2 b2 + b1
            if ((lane & 1) == 0 && comp < 12 && v1 != 0.0) atomicAdd(&sm.gacc[k0][comp], v1);
        }
    }
    __syncthreads();
    if (tid < take * 12 && p.gmvp) {
        const int s = tid / 12, k = tid - s * 12;
        const double v = sm.gacc[s][k];
        // rows x (0), y (1), w (3) of d loss / d mvp; the z row carries no gradient
        if (v != 0.0) atomicAdd(p.gmvp + ((size_t)c.item * p.L + sm.slot[s].link) * 16 + (k < 8 ? k : k + 4), v);
    }
    EHB_STAT_ADD(10, clock64() - tF);
}

// MODE: 0 = fused / antialiased forward (needs the masks), 1 = operator backward (g comes from the caller).
// REFKIND: 0 none, 1 f32, 2 u8, 3 registered bits.  BWD: the backward follows the forward in the same pass.
template <int MODE, int REFKIND, bool BWD>
__global__ void __launch_bounds__(EHB_TTHREADS, EHB_TMIN_BLOCKS) ehb_k_tiles(const __grid_constant__ EhbRobot rb,
                                                               const __grid_constant__ EhbParams p)
{
    ehb_pdl_enter();
    EHB_TL_START(tl0);
    EHB_MARK(p, 10);
    __shared__ EhbTileSm sm;
    constexpr bool NEEDAA = MODE == 0;
    constexpr int OW = EHB_T + ((NEEDAA && BWD) ? 1 : 0);   // out region whose S is needed (33 when g is needed on it)
    constexpr bool OEXT = NEEDAA && BWD;
    EhbTileCtx c;
    c.tid = threadIdx.x; c.lane = c.tid & 31; c.warp = c.tid >> 5;
    const int tid = c.tid, lane = c.lane, warp = c.warp;
    const int H = p.H, W = p.W;
    const unsigned nHeavy = p.ctr->nTiles, nEntries = nHeavy + p.ctr->nLight;
    const uint32_t linkMask = p.L >= 32 ? 0xFFFFFFFFu : ((1u << p.L) - 1u);
    const bool tma = NEEDAA && p.masks != nullptr && p.useTma;
    float* stage = sm.mbuf[0];
    bool storePending = false;                           // thread 0: a TMA store may still be reading the staging tile
    if (tid == 0) sm.slab = -1;

    for (unsigned e = blockIdx.x; e < nEntries; e += gridDim.x) {
        EHB_STAT_T(tTile);
#ifdef EHB_STATS
        if (tid == 0) for (int k = 0; k < 8; k++) sm.stat[k] = 0;
#endif
        const uint32_t wid = ehb_list_at(p, e, nHeavy);
        c.item = (int)(wid / (uint32_t)p.ntiles); c.tile = (int)(wid - (uint32_t)c.item * (uint32_t)p.ntiles);
        const int tyy = c.tile / p.ntx;
        c.x0 = (c.tile - tyy * p.ntx) * EHB_T; c.y0 = tyy * EHB_T;
        c.rx0 = c.x0 - 1; c.ry0 = c.y0 - 1;              // window = tile + 1 low / 2 high halo pixels (35 x 35), all modes
        const int item = c.item, x0 = c.x0, y0 = c.y0;
        const size_t ibase = (size_t)item * H * W;
        __syncthreads();                                 // the previous tile is done with shared memory
        if (tid == 0 && storePending) { ehb_bulk_wait_read(); storePending = false; }
        // one round trip fetches everything that depends only on the tile: its link bits, the item's depth planes, the
        // reference bits of its out region and its reference count
        if (tid == 0) {
            sm.bits = p.touch[wid] & linkMask;
            sm.refCnt = REFKIND == 3 ? __ldg(p.refCnt + (size_t)item * p.ntiles + c.tile) : 0u;
        }
        if (tid < p.L) sm.planes[tid] = p.plane[(size_t)item * p.L + tid];
        if (REFKIND == 3 && tid >= 32 && tid < 32 + 2 * EHB_MROWS) {
            const int k = tid - 32, qy = k >> 1, wx = (x0 >> 5) + (k & 1), py = y0 + qy;
            sm.rbits[qy][k & 1] = (py < H && wx < p.ntx) ? __ldg(p.refBits + ((size_t)item * H + py) * p.ntx + wx) : 0u;
        }
        if (MODE == 1) {                                 // g = dL/dmask comes from the caller
            for (int qy = warp; qy < EHB_MROWS; qy += EHB_TWARPS) {
                const int py = y0 + qy;
#pragma unroll
                for (int part = 0; part < 2; part++) {
                    const int qx = part ? EHB_T : lane, px = x0 + qx;
                    if (part && lane != 0) break;
                    sm.S[qy * EHB_MW + qx] = (px < W && py < H) ? __ldg(p.dy + ibase + (size_t)(H - 1 - py) * W + px) : 0.f;
                }
            }
        }
        __syncthreads();
        c.bits = sm.bits;
        c.nl = __popc(c.bits);
        const int nl = c.nl;
        EHB_STAT_ADD(0, 1); EHB_STAT_ADD(1, nl); EHB_STAT_ADD(2, nl == 0);
        // ---- a listed tile that no triangle reaches: a zero tile (registered reference: its loss is part of refTotal) ------
        if (nl == 0 && (REFKIND == 0 || REFKIND == 3)) {
            if (NEEDAA && p.masks) {
                if (tma) {
                    for (int i = tid; i < EHB_T * EHB_T / 4; i += EHB_TTHREADS) reinterpret_cast<float4*>(stage)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    ehb_fence_proxy_async();
                    __syncthreads();
                    if (tid == 0) {
                        const int row0 = H - EHB_T - y0;
                        ehb_tma_store_3d(row0 < 0 ? &p.tmMaskTop : &p.tmMask, stage, x0, max(row0, 0), item);
                        ehb_bulk_commit();
                        storePending = true;
                    }
                } else {
                    for (int i = tid; i < EHB_T * EHB_T; i += EHB_TTHREADS) {
                        const int px = x0 + (i & 31), py = y0 + (i >> 5);
                        if (px < W && py < H) p.masks[ibase + (size_t)(H - 1 - py) * W + px] = 0.f;
                    }
                }
            }
            EHB_STAT_ADD(11, clock64() - tTile);
            continue;
        }
        // ---- rounds over the tile's links (one round for the tiles of a robot arm) -------------------------------------
        int lNext = 0, rounds = 0, take = 0, nPairsRound = 0;
        EhbPairs pr;
        pr.alpha = sm.alpha; pr.ptri = sm.ptri; pr.pk = sm.pk; pr.pslot = sm.pslot;
        while (lNext < nl) {
            ehb_tile_round<NEEDAA, OW>(rb, p, sm, c, lNext, take, nPairsRound, pr);
            if (NEEDAA) {
                ehb_tile_masks<OW>(p, sm, c, pr, take, nPairsRound, rounds == 0);
            } else {
                ehb_tile_backward(rb, p, sm, c, pr, take, nPairsRound);          // operator backward: g is already there
            }
            lNext += take; rounds++;
        }
        EHB_STAT_ADD(4, rounds > 1);
        if (!NEEDAA) { EHB_STAT_ADD(11, clock64() - tTile); continue; }
        // ================================ E: compose, loss, dL/dsum ================================
        // warp w takes the rows qy = w, w + 4, ...; lane = column; S holds the sum of the link masks
        __syncthreads();
        EHB_STAT_T(tE);
        double lacc = 0.0;
        auto pixel = [&](int qy, int qx, uint32_t refBit, bool interior) {
            const int px = x0 + qx, py = y0 + qy;
            const bool inImg = py < H && px < W;
            const float s = nl > 0 ? sm.S[qy * EHB_MW + qx] : 0.f;
            const float Sv = (p.clamp && s > 1.f) ? 1.f : s;
            float gv = 0.f;
            if (REFKIND && inImg) {
                float rf;
                if (REFKIND == 3) rf = refBit ? 1.f : 0.f;
                else {
                    const size_t o = ibase + (size_t)(H - 1 - py) * W + px;
                    rf = REFKIND == 1 ? __ldg(p.ref + o) : (__ldg(p.ref_u8 + o) ? 1.f : 0.f);
                }
                const float diff = Sv - rf;
                if (interior) lacc += (double)(diff * diff);
                gv = (!p.clamp || s <= 1.f) ? (2.f * diff) * p.invB : 0.f;
            }
            if (OEXT) sm.S[qy * EHB_MW + qx] = gv;
            return Sv;
        };
        float keep[(EHB_T + EHB_TWARPS - 1) / EHB_TWARPS];   // this lane's composed values of the interior rows
#pragma unroll
        for (int j = 0; j < (OW + EHB_TWARPS - 1) / EHB_TWARPS; j++) {
            const int qy = warp + j * EHB_TWARPS;
            if (qy < OW) {
                const float Sv = pixel(qy, lane, REFKIND == 3 ? (sm.rbits[qy][0] >> lane) & 1u : 0u, qy < EHB_T);
                if (j < (EHB_T + EHB_TWARPS - 1) / EHB_TWARPS) keep[j] = Sv;
            }
        }
        if (OEXT && warp == EHB_TWARPS - 1) {            // column 32 of the out region (33 pixels): g only
            pixel(lane, EHB_T, REFKIND == 3 ? sm.rbits[lane][1] & 1u : 0u, false);
            if (lane == 0) pixel(EHB_T, EHB_T, REFKIND == 3 ? sm.rbits[EHB_T][1] & 1u : 0u, false);
        }
        if (REFKIND && p.loss) {
            lacc = ehb_warp_sum(lacc);
            if (lane == 0) sm.lsum[warp] = lacc;
        }
        if (p.masks) {
            // (the staging tile aliases the mask buffer, which is dead: S is complete)
#pragma unroll
            for (int j = 0; j < (EHB_T + EHB_TWARPS - 1) / EHB_TWARPS; j++) {
                const int qy = warp + j * EHB_TWARPS, py = y0 + qy;
                if (qy < EHB_T && py < H) {
                    if (tma) stage[(H - 1 - py - max(H - EHB_T - y0, 0)) * EHB_T + lane] = keep[j];          // image rows run downwards
                    else if (x0 + lane < W) p.masks[ibase + (size_t)(H - 1 - py) * W + x0 + lane] = keep[j];
                }
            }
            if (tma) ehb_fence_proxy_async();
        }
        __syncthreads();
        if (tid == 0) {
            if (tma) {
                // the tile's image rows: H - 32 - y0 .. H - 1 - y0; the top row of tiles of an image whose height is not a
                // multiple of 32 starts at image row 0 and uses the shorter box
                const int row0 = H - EHB_T - y0;
                ehb_tma_store_3d(row0 < 0 ? &p.tmMaskTop : &p.tmMask, stage, x0, max(row0, 0), item);
                ehb_bulk_commit();
                storePending = true;
            }
            if (REFKIND && p.loss) {
                double t = 0.0;
                for (int w = 0; w < EHB_TWARPS; w++) t += sm.lsum[w];
                if (REFKIND == 3) t -= (double)sm.refCnt;   // loss[item] starts at sum(ref)
                if (t != 0.0) atomicAdd(&p.loss[item], t);
            }
        }
        EHB_STAT_ADD(9, clock64() - tE);
        if (!BWD || nl == 0) { EHB_STAT_ADD(11, clock64() - tTile); continue; }
        if (rounds == 1) {
            ehb_tile_backward(rb, p, sm, c, pr, take, nPairsRound);          // the pairs of the forward are still resident
        } else {
            // g is known now: rebuild each round's pairs for the backward
            lNext = 0;
            while (lNext < nl) {
                ehb_tile_round<NEEDAA, OW>(rb, p, sm, c, lNext, take, nPairsRound, pr);
                ehb_tile_backward(rb, p, sm, c, pr, take, nPairsRound);
                lNext += take;
            }
        }
        EHB_STAT_ADD(11, clock64() - tTile);
#ifdef EHB_STATS
        if (tid == 0 && p.dbgbuf && e < 4096u) {
            for (int k = 0; k < 7; k++) p.dbgbuf[(size_t)e * 8 + k] = (unsigned long long)sm.stat[k];
            p.dbgbuf[(size_t)e * 8 + 7] = (unsigned long long)nl | ((unsigned long long)sm.stat[7] << 8);
        }
#endif
    }
    if (tid == 0 && storePending) ehb_bulk_wait_read();
    // The tiles no link touches (registered reference masks: their loss is part of refTotal, nothing is read): mask := 0, a
    // warp per tile with 16-B streaming stores, by every CTA of this launch once its own tiles are done -- the CTAs beyond
    // the tile list start with it.  31 MB of HBM writes for ten views that overlap the latency-bound tile work.  (Measured
    // alternatives: a bulk tensor store (UTMASTG) of a zero tile per untouched tile costs 0.2 us per tile and SM, 18 - 33 us
    // per pass; as spare CTAs of the raster launch the fill held a quarter of its CTA slots for 6 - 10 us; as CTAs of
    // k_front it was the longest part of that launch, 13 us.)
    if (NEEDAA && p.fillEmpty) {
        const int nEmpty = (int)p.ctr->nEmpty;
        for (int i = (int)blockIdx.x * EHB_TWARPS + warp; i < nEmpty; i += (int)gridDim.x * EHB_TWARPS) {
            const int wid = (int)p.emptyList[i];
            const int it = wid / p.ntiles, tile = wid - it * p.ntiles;
            ehb_stream_empty_tile(p, it, tile % p.ntx, tile / p.ntx, lane);
        }
    }
    if (tid == 0) EHB_TL_STOP(p, 4, blockIdx.x, tl0);
}

// the instantiation of a pass
typedef void (*EhbTilesKernel)(const EhbRobot, const EhbParams);
inline EhbTilesKernel ehb_tiles_kernel(int mode, int refKind, bool bwd)
{
    if (mode == EHB_MODE_AA_BWD) return ehb_k_tiles<1, 0, true>;
    if (mode != EHB_MODE_FUSED) return ehb_k_tiles<0, 0, false>;
    switch (refKind * 2 + (bwd ? 1 : 0)) {
    case 0: case 1: return ehb_k_tiles<0, 0, false>;
    case 2: return ehb_k_tiles<0, 1, false>;
    case 3: return ehb_k_tiles<0, 1, true>;
    case 4: return ehb_k_tiles<0, 2, false>;
    case 5: return ehb_k_tiles<0, 2, true>;
    case 6: return ehb_k_tiles<0, 3, false>;
    default: return ehb_k_tiles<0, 3, true>;
    }
}

}  // namespace EHB_TNS
