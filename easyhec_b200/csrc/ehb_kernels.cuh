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
//              small are deferred (record parked in global memory, bbox cut into 64x32 units), and a batch with many
//              rows keeps only its first groups (records + row prefix parked): both go to k_raster_big.
//              Spare CTAs of the same launch stream the empty tiles (mask = 0, loss += ref^2, float4): the HBM-bound
//              part of the frame overlaps the instruction-bound part instead of preceding it.
//   k_raster_big : warps stride over the units (bounded work each, spread over the whole chip); spare CTAs build the
//              job list of the image-space stage from the touch bitmap.
// The image-space stage (antialias, compose, loss, backward) is in ehb_tiles.cuh.
// No intermediate image (rast, colour, antialias work queue, clip-space vertex buffer) is ever written.
#pragma once
#include "ehb_device.cuh"
#include "ehb_tma.cuh"

#define EHB_T 32            // tile interior
#define EHB_RS 35           // window row stride = T + max halo (1 low, 2 high)
#define EHB_NP (EHB_RS * EHB_RS)

enum { EHB_MODE_FUSED = 0, EHB_MODE_AA_FWD = 1, EHB_MODE_AA_BWD = 2, EHB_MODE_UNION = 3 };

struct EhbPlane {            // depth plane of one (item, link): pixels [x0, x0+w) x [y0, y0+h), GL rows
    int x0, y0, w, h;
    long long off;           // first element in the plane pool
    long long boff;          // first 64-bit word of the plane's coverage bits in the bit pool (h rows of ceil(w / 64) words)
};

struct __align__(8) EhbUnit { uint32_t rec; unsigned short dx0, dy0; };   // a 64 x 32 pixel window of a deferred triangle's bbox

struct __align__(128) EhbCounters {
    // line 0: pass bookkeeping (one writer at a time, or a few hundred atomics per pass)
    unsigned long long planeCursor;
    unsigned long long nNeedClip;
    unsigned int nTiles;     // tiles touched by >= 2 link bboxes: listed from the front of tileList (served first) ...
    unsigned int nLight;     // ... the others from the back
    unsigned int workCursor;
    unsigned int nEmpty;     // tiles no link touches: streamed (mask = 0, loss += ref^2) by spare CTAs of the raster launch
    unsigned int vertexDone; // CTAs of k_table that have finished: the last one allocates the planes
    unsigned int flags;      // 1: plane pool too small (results invalid, grow and rerun), 2: triangles need clipping
    unsigned int nBatchHeavy; // visible 32-triangle batches with a large screen footprint: listed from the front of batchList ...
    unsigned int nBatchLight; // ... the other visible ones from the back (k_front writes, k_raster reads, k_raster_big resets)
    unsigned long long bitCursor;   // words of the bit pool in use
    unsigned long long prevCursor, prevBitCursor;   // what the previous pass on this scratch used: the part k_front clears
    unsigned int pad0[14];
    // line 1: set by the last table CTA of k_front when every plane's bounding box is written and the list counters are
    // reset; the tile-list CTAs of the same launch wait for it (k_raster resets it)
    unsigned int tableReady;
    unsigned int pad1[31];
    unsigned int pad2[32];
    // lines 4..35: the queues of deferred work, split into EHB_NQ sub-queues with their counters on separate lines --
    // thousands of same-address atomics per pass would serialise on one L2 line (measured: a third of k_raster's
    // warp-time waiting for them); a warp uses the sub-queue of its index
    struct Q {
        unsigned int nBigRec;    // deferred (not small) triangles: records parked in global memory ...
        unsigned int nUnits;     // ... and cut into bounded units that k_raster_big spreads over the whole chip
        unsigned int nBatchBlk;  // batches with many rows: records parked, the rows beyond the inline share become units too
        unsigned int take;       // k_raster_big: units of this sub-queue handed out so far
        unsigned int pad[28];
    } q[32];
    // pair pool of the image-space stage: slabs handed to the rare tile whose pairs do not fit shared memory
    unsigned int slabCursor;
    unsigned int pad3[31];
    unsigned long long dbg[16];   // EHB_TIMING builds: cycles per phase of k_tiles (thread 0 of every CTA)
};
static_assert(sizeof(EhbCounters) % 128 == 0, "counter lines");

struct EhbParams {
    // masks as a [items][H][W] f32 tensor (valid when useTma): box 32 x 32 x 1, and box 32 x (H % 32) x 1 for the top row
    // of tiles when H is not a multiple of 32 -- a tile store may stick out of the tensor on the high side (clipped by the
    // hardware) but must not start at a negative coordinate (measured: illegal instruction), so the partial tiles at the
    // top of the image get their own, shorter box that starts at image row 0
    CUtensorMap tmMask, tmMaskTop;
    int useTma;
    int fillEmpty;           // k_tiles zero-fills the tiles no link touches (no reference to read there)
    int H, W, ntx, nty, ntiles;
    int items, L, Lp, Ftot, Vtot;   // Lp = planes per item: L (per-link visibility) or 1 (packed robot)
    int hlo, hhi;
    int smallArea, inlineGroups;   // per-pass knobs of the rasterizer (see EHB_SMALL_AREA_SERIAL / _INFLIGHT)
    uint32_t* batchList;     // [items * chunks]  the batches that survive the frustum test of their AABB (k_front)
    float heavyArea;         // screen footprint (px^2) of a batch's AABB from which it is listed in front
    int mode, rule, do_bwd, clamp;
    float invB;
    float xs, xo, ys, yo;    // NDC of a pixel centre: x = xs * px + xo, y = ys * py + yo (2/W, 1/W - 1, 2/H, 1/H - 1 in fp32)
    const float* mvp;        // [items, L, 16]
    float4* vclip;           // [items, Vtot]  clip-space position of every vertex (written by k_front)
    int2* vsnap;             // [items, Vtot]  snapped screen position (1/16 px), x = INT_MIN when not drawable
    EhbPlane* plane;         // [items, Lp]
    unsigned long long* pool;
    unsigned long long poolCap;
    unsigned long long* bits;    // coverage bit per pixel of every plane (per-link visibility only): what the image-space stage
    unsigned long long bitCap;   //   reads instead of the 64 times larger depth planes; NULL in the packed-robot mode
    struct EhbBitsRec* bigBits;  // [bigCap * EHB_NQ]  bit-plane address of every parked record
    uint32_t* tileList;      // [items * ntiles]
    uint32_t* emptyList;     // [items * ntiles]
    uint32_t* touch;         // [items * ntiles]  bit l: a triangle of link l reaches into this tile's window
    struct EhbRec* bigRec;   // [bigCap]
    EhbUnit* units;          // [unitCap]
    int bigCap, unitCap;     // per sub-queue (the arrays hold EHB_NQ times as many)
    uint32_t* batchBlk;      // [batchCap][EHB_BLK_WORDS] parked batches: 32 records (transposed) + row prefix + flags
    int batchCap;            // per sub-queue
    EhbCounters* ctr;
    const float* ref;        // [items, H, W]  FUSED
    const uint8_t* ref_u8;   // same, as bytes (either ref or ref_u8)
    float* masks;            // [items, H, W]  FUSED / AA_FWD
    double* loss;            // [items]
    double* gmvp;            // [items, L, 16]
    float* gpos;             // [V, 4] (AA_BWD, single link) or NULL
    const float* dy;         // [items, H, W]  AA_BWD
    uint8_t* out_u8;         // [items, H, W]  UNION
    // reference masks registered once (ehb_ref_register): one bit per pixel, GL rows, a 32-bit word per (row, tile column);
    // refCnt = set bits of every tile's interior, refTotal = set bits of the item.  loss[item] starts at refTotal and every
    // listed tile adds (its loss - its refCnt): the tiles no link touches are never read.
    unsigned char* pairPool; // [nSlabs][EHB_SLAB_BYTES]  pair lists that do not fit a CTA's shared memory (ehb_tiles.cuh)
    int nSlabs;
    const uint32_t* refBits; // [items, H, ntx]
    const uint32_t* refCnt;  // [items, ntiles]
    const unsigned long long* refTotal;   // [items]
    unsigned long long* dbgbuf;   // unused (kept for the developer ABI)
    unsigned int* hostFlags;      // mapped pinned host memory: word b is set when flag bit b is raised (ehb_ctx_poll reads it
                                  // without synchronising)
};

// Coverage bits of a plane: bit (px - x0) of row (py - y0); a row is ceil(w / 64) words.  `base` folds the plane's origin in:
// bit index in the pool = base + py * pitch + px.
struct EhbBits { unsigned long long* w; long long base; int pitch; };
struct __align__(16) EhbBitsRec { long long base; int pitch; int pad; };
__device__ __forceinline__ EhbBits ehb_bits_of(unsigned long long* words, const EhbPlane& pl)
{
    EhbBits b;
    b.w = words;
    b.pitch = ((pl.w + 63) >> 6) << 6;
    b.base = pl.boff * 64 - (long long)pl.y0 * b.pitch - pl.x0;
    return b;
}
// set the bits of `len` >= 1 pixels of one row, starting at (px, py): one RED.OR per 64-bit word the run touches (one for
// nearly every run: the runs of small triangles are a few pixels long)
__device__ __forceinline__ void ehb_bits_set(const EhbBits& b, int px, int py, int len)
{
    const long long pos = b.base + (long long)(py * b.pitch + px);   // (py * pitch + px < 2^27)
    unsigned long long* wp = b.w + (pos >> 6);
    const int bo = (int)((unsigned)pos & 63u);
    const int n0 = min(len, 64 - bo);
    atomicOr(wp, (n0 == 64 ? ~0ull : ((1ull << n0) - 1ull)) << bo);
    int rem = len - n0;
    while (rem > 0) {                                                // (rare: the run crosses a word boundary)
        wp++;
        atomicOr(wp, rem >= 64 ? ~0ull : ((1ull << rem) - 1ull));
        rem -= 64;
    }
}

// Raise sticky status bits: in the pass's counters (read by ehb_ctx_status) and in host-visible memory (ehb_ctx_poll).
__device__ __forceinline__ void ehb_raise(const EhbParams& p, unsigned bits)
{
    atomicOr(&p.ctr->flags, bits);
    if (p.hostFlags) {
        for (unsigned b = 0; b < 3; b++)
            if (bits & (1u << b)) *reinterpret_cast<volatile unsigned int*>(p.hostFlags + b) = 1u;
    }
}

// EHB_TIMELINE builds (developer tool, tools/timeline.py): start / end of every warp or CTA of a pass on the global timer,
// region k of the debug buffer (ehb_ctx_debug_buffer): 0 table, 1 front, 2 raster (per warp), 3 raster_big (per warp),
// 4 tiles (per CTA), 5 raster's stream CTAs (per warp)
#define EHB_TL_N 16384
#ifdef EHB_TIMELINE
__device__ __forceinline__ unsigned long long ehb_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned ehb_smid() { unsigned s; asm("mov.u32 %0, %%smid;" : "=r"(s)); return s; }
#define EHB_TL_START(var) const unsigned long long var = ehb_gtime()
#define EHB_TL_STOP(p, k, idx, var) do { if ((p).dbgbuf && (unsigned)(idx) < EHB_TL_N) { (p).dbgbuf[((size_t)(k) * EHB_TL_N + (idx)) * 2] = (var); \
    (p).dbgbuf[((size_t)(k) * EHB_TL_N + (idx)) * 2 + 1] = (ehb_gtime() << 8) | ehb_smid(); } } while (0)
#else
#define EHB_TL_START(var)
#define EHB_TL_STOP(p, k, idx, var)
#endif

#ifdef EHB_MARKS
#define EHB_MARK(p, i) do { if ((p).hostFlags) *reinterpret_cast<volatile unsigned int*>((p).hostFlags + (i)) = 1u; } while (0)
#else
#define EHB_MARK(p, i)
#endif
// Two knobs of the rasterizer are per pass (EhbParams): smallArea -- triangles whose clipped bbox has more candidate
// samples are deferred to k_raster_big -- and inlineGroups -- groups of 32 rows a warp of k_raster draws itself before it
// hands the rest of a heavy batch on.  A pass that runs alone wants its tails short (96 samples, 2 groups: the work
// spreads over the chip early); passes in flight on the slots overlap each other's tails and want the fewest instructions
// (256 samples, 4 groups: less parking and re-fetching, +8 % frames/s).  Measured on the B200, see DESIGN.md.
#define EHB_SMALL_AREA_SERIAL 96
#define EHB_INLINE_SERIAL 2
#define EHB_SMALL_AREA_INFLIGHT 256
#define EHB_INLINE_INFLIGHT 4
#define EHB_NQ 32
#define EHB_UNIT_W 64
#define EHB_UNIT_H 32

// Programmatic dependent launch (sm_90+): the kernels of a pass are chained with the programmatic-stream-serialization
// launch attribute.  Each kernel first lets its successor be scheduled (its CTAs take free SM slots during this kernel's
// tail and run their prologue), then waits until its predecessor has completed and flushed.
__device__ __forceinline__ void ehb_pdl_enter()
{
#ifdef EHB_PDL
#if EHB_PDL == 2
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
    // EHB_PDL == 1: no early trigger -- the dependent grid is released when the CTAs of this one exit, so it never
    // competes with this grid's own unscheduled CTAs for SM slots; only its launch latency is hidden
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

// ------------------------------------------------------------------------------------------------ empty tiles
// A tile no link touches: mask = 0 and loss += sum ref^2, streamed by one warp.  This is the HBM-bound 85 % of a frame,
// so every load of the tile is issued before the first one is consumed (8 x 512 B in flight per warp for an f32
// reference, the whole 1 KB tile in two 16-B loads per lane for a u8 one); the zero stores follow.
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
            const int pyb = y0 + (lane >> 3);
            // image rows run downwards while py runs upwards: o(it) = o0 - it * 4 * W
            const size_t o0 = ibase + (size_t)(H - 1 - pyb) * W + cx;
            const size_t st = (size_t)4 * W;
            const bool colOk = cx < W;
            if (p.ref) {
                float4 r[8];
#pragma unroll
                for (int it = 0; it < 8; it++) {
                    r[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (colOk && pyb + it * 4 < H) r[it] = __ldg(reinterpret_cast<const float4*>(p.ref + (o0 - it * st)));
                }
#pragma unroll
                for (int it = 0; it < 8; it++)
                    acc += (double)(r[it].x * r[it].x) + (double)(r[it].y * r[it].y) + (double)(r[it].z * r[it].z) + (double)(r[it].w * r[it].w);
            } else if (p.ref_u8) {
                if ((W & 15) == 0 && x0 + EHB_T <= W && (((uintptr_t)p.ref_u8) & 15) == 0) {
                    // one lane per row of the tile: its 32 bytes are two aligned 16-B words
                    const int py = y0 + lane;
                    if (py < H) {
                        const uint4* src = reinterpret_cast<const uint4*>(p.ref_u8 + ibase + (size_t)(H - 1 - py) * W + x0);
                        const uint4 a = __ldg(src), b = __ldg(src + 1);
                        const uint32_t w8[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                        int n = 0;
#pragma unroll
                        for (int k = 0; k < 8; k++) {
                            // bytes != 0, counted four at a time: a byte is non-zero iff (b | (b + 0x7f)) has its top bit set
                            const uint32_t v = w8[k];
                            const uint32_t nz = ((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v;
                            n += __popc(nz & 0x80808080u);
                        }
                        acc += (double)n;
                    }
                } else {
                    uchar4 r[8];
#pragma unroll
                    for (int it = 0; it < 8; it++) {
                        r[it] = make_uchar4(0, 0, 0, 0);
                        if (colOk && pyb + it * 4 < H) r[it] = __ldg(reinterpret_cast<const uchar4*>(p.ref_u8 + (o0 - it * st)));
                    }
#pragma unroll
                    for (int it = 0; it < 8; it++) acc += (double)((r[it].x != 0) + (r[it].y != 0) + (r[it].z != 0) + (r[it].w != 0));
                }
            }
            if (p.masks && colOk) {
#pragma unroll
                for (int it = 0; it < 8; it++)
                    if (pyb + it * 4 < H) __stcs(reinterpret_cast<float4*>(p.masks + (o0 - it * st)), make_float4(0.f, 0.f, 0.f, 0.f));
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
// ONE launch in front of the rasterizer, five kinds of CTAs (block ranges, in this order):
//   (T) table     one warp per depth plane: conservative screen bounding box of the link from the projected corners of its
//                 (up to 32) object-space chunk AABBs -- 256 point transforms instead of a reduction over all vertices.  The
//                 last table CTA to finish (atomic ticket) allocates the planes of the pass in the pool by a block-wide
//                 prefix sum (deterministic); before that -- as soon as every box is written -- it raises ctr->tableReady
//                 (the first table CTA has reset the list counters by then).
//   (V) vertices  one thread per (item, vertex): transform and snap ONCE (a vertex is shared by ~6 triangles), keep the
//                 clip-space and snapped positions (a few MB, L2-resident)
//   (B) batches   frustum test of every 32-triangle batch's object-space AABB (8 corners, one per lane): the batches that
//                 survive are listed -- large screen footprints from the front, the rest from the back -- so that k_raster
//                 runs only warps that have something to draw, the long ones first
//   (C) clear     what the previous pass used of the plane and bit pools := EMPTY / 0, link bitmaps := 0
//   (L) tiles     one lane per tile: tiles that some link's bbox touches -> tile list (several links first), the others ->
//                 empty-tile list.  These CTAs wait for tableReady (the table CTAs have the lowest block indices: they are
//                 resident before any waiter).
// Only L depends on the table; V, B and C run beside it.  (As two launches -- table, then the rest -- the pass paid one more launch gap and the table's 4 us with
// 9 CTAs on the chip.)
__device__ __forceinline__ void ehb_wait_table(const EhbParams& p)
{
    if (threadIdx.x == 0) {
        unsigned v;
        EHB_MARK(p, 11);
        for (;;) {
            asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(&p.ctr->tableReady) : "memory");
            if (v) break;
            __nanosleep(64);
        }
        EHB_MARK(p, 7);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) ehb_k_front(const __grid_constant__ EhbRobot rb, const __grid_constant__ EhbParams p,
                                                   int tableBlocks, int vchunks, int batchBlocks, int clearBlocks, int chunks,
                                                   int blockBase)
{
    ehb_pdl_enter();
    EHB_TL_START(tl0);
    const int lane = threadIdx.x & 31;
    int blk = (int)blockIdx.x + blockBase;   // (blockBase != 0: the launch without the table CTAs, developer switch EHB_FRONT_SPLIT)
    // ---------------------------------------------------------------- (T) table
    if (blk < tableBlocks) {
        EHB_MARK(p, 4);
        // the first table CTA resets the counters the tile-list CTAs and the rasterizer add to (nobody touches them before the
        // table is ready, the previous pass is complete): off the critical path of the last CTA's hand-off
        if (blk == 0) {
            if (threadIdx.x == 0) { p.ctr->nTiles = 0u; p.ctr->nLight = 0u; p.ctr->nEmpty = 0u; p.ctr->workCursor = 0u; p.ctr->slabCursor = 0u; }
            if (threadIdx.x < EHB_NQ) {
                p.ctr->q[threadIdx.x].nBigRec = 0u; p.ctr->q[threadIdx.x].nUnits = 0u; p.ctr->q[threadIdx.x].nBatchBlk = 0u;
                p.ctr->q[threadIdx.x].take = 0u;
            }
        }
        const int wid = (blk * (int)blockDim.x + (int)threadIdx.x) >> 5;
        // outputs that the later kernels accumulate into
        for (int i = blk * blockDim.x + threadIdx.x; i < p.items; i += tableBlocks * blockDim.x)
            if (p.loss) p.loss[i] = p.refTotal ? (double)p.refTotal[i] : 0.0;
        if (p.gmvp)
            for (int i = blk * blockDim.x + threadIdx.x; i < p.items * p.L * 16; i += tableBlocks * blockDim.x) p.gmvp[i] = 0.0;
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
                pl.x0 = pl.y0 = pl.w = pl.h = 0; pl.off = 0; pl.boff = 0;
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
        // the last table CTA to finish sees every bounding box: it allocates the planes of this pass
        __shared__ unsigned s_last;
        __shared__ unsigned long long s_wsum[8], s_base, s_wsumb[8], s_baseb;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = atomicAdd(&p.ctr->vertexDone, 1u) == (unsigned)tableBlocks - 1u;
        __syncthreads();
        if (!s_last) { if (threadIdx.x == 0) EHB_TL_STOP(p, 0, blockIdx.x, tl0); return; }
        __threadfence();
        EHB_MARK(p, 5);
        // the tile-list CTAs need the bounding boxes and the reset counters (every table CTA fenced its writes before its
        // ticket), not the allocation that follows (only the kernels after this launch read it): they are released now
        if (threadIdx.x == 0) {
            asm volatile("st.release.gpu.u32 [%0], %1;" ::"l"(&p.ctr->tableReady), "r"(1u) : "memory");
            s_base = 0ull; s_baseb = 0ull;
            p.ctr->vertexDone = 0u;
        }
        __syncthreads();
        const int warp = threadIdx.x >> 5;
        for (int i0 = 0; i0 < p.items * p.Lp; i0 += blockDim.x) {      // block-wide exclusive prefix sums: plane areas, bit words
            const int i = i0 + threadIdx.x;
            EhbPlane pl;
            pl.x0 = pl.y0 = pl.w = pl.h = 0; pl.off = 0; pl.boff = 0;
            if (i < p.items * p.Lp) {
                const int4 a = __ldcg(reinterpret_cast<const int4*>(&p.plane[i]));   // written by other CTAs of this launch: read at L2
                pl.x0 = a.x; pl.y0 = a.y; pl.w = a.z; pl.h = a.w;
            }
            const unsigned long long area = (unsigned long long)pl.w * (unsigned long long)pl.h;
            const unsigned long long words = p.bits ? (unsigned long long)((pl.w + 63) >> 6) * (unsigned long long)pl.h : 0ull;
            unsigned long long inc = area, incb = words;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o), vb = __shfl_up_sync(0xffffffffu, incb, o);
                if (lane >= o) { inc += v; incb += vb; }
            }
            if (lane == 31) { s_wsum[warp] = inc; s_wsumb[warp] = incb; }
            __syncthreads();
            unsigned long long before = s_base, beforeb = s_baseb;
            for (int w = 0; w < warp; w++) { before += s_wsum[w]; beforeb += s_wsumb[w]; }
            const unsigned long long off = before + inc - area, offb = beforeb + incb - words;
            if (pl.w > 0) {
                if (off + area <= p.poolCap && offb + words <= p.bitCap) { pl.off = (long long)off; pl.boff = (long long)offb; }
                else { pl.w = pl.h = 0; ehb_raise(p, 1u); }
                p.plane[i] = pl;
            }
            __syncthreads();
            if (threadIdx.x == blockDim.x - 1) { s_base = before + inc; s_baseb = beforeb + incb; }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            p.ctr->planeCursor = min(s_base, p.poolCap);
            p.ctr->bitCursor = min(s_baseb, p.bitCap);
            EHB_MARK(p, 6);
            EHB_TL_STOP(p, 0, blockIdx.x, tl0);
        }
        return;
    }
    blk -= tableBlocks;
    // ---------------------------------------------------------------- (V) vertices
    const int vertexBlocks = vchunks * p.items;
    if (blk < vertexBlocks) {
        const int item = blk / vchunks;
        const int g = (blk - item * vchunks) * blockDim.x + threadIdx.x;
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
        if (threadIdx.x == 0) EHB_TL_STOP(p, 1, blockIdx.x, tl0);
        return;
    }
    blk -= vertexBlocks;
    // ---------------------------------------------------------------- (B) visible batches
    if (blk < batchBlocks) {
        // 8 lanes per batch (one per corner of its AABB), 32 batches per CTA.  When every corner is in front of the camera
        // and all of them lie beyond one side of the screen (two pixels of margin for snapping and rounding), every vertex
        // of the batch does too (projection keeps convex hulls when w > 0), so each of its triangles would fail the
        // per-triangle tests: the batch is not listed.
        const int total = chunks * p.items;
        const int b = blk * 32 + (int)(threadIdx.x >> 3);
        const int k = threadIdx.x & 7;
        const unsigned gm = 0xFFu << (lane & 24);             // the lanes of this batch
        bool front = false, oxr = false, oxl = false, oyt = false, oyb = false, emptyBox = true;
        float u = 0.f, v = 0.f;
        if (b < total) {
            const int item = b / chunks, bl = b - item * chunks;
            const int link = ehb_find_link(rb.boff, rb.L, bl);
            const float4* fb = rb.link[link].fboxes + 2 * (bl - rb.boff[link]);
            const float4 lo = __ldg(fb), hi = __ldg(fb + 1);
            float m[16], c[4];
            ehb_load_mvp(p.mvp + ((size_t)item * p.L + link) * 16, m);
            ehb_xform(make_float4((k & 1) ? hi.x : lo.x, (k & 2) ? hi.y : lo.y, (k & 4) ? hi.z : lo.z, 1.f), m, c);
            front = c[3] > 1e-6f;
            const float mx = 4.f / (float)p.W, my = 4.f / (float)p.H;
            oxr = c[0] - (1.f + mx) * c[3] > 0.f; oxl = -c[0] - (1.f + mx) * c[3] > 0.f;   // beyond the right / left side
            oyt = c[1] - (1.f + my) * c[3] > 0.f; oyb = -c[1] - (1.f + my) * c[3] > 0.f;
            emptyBox = lo.x > hi.x;
            if (front) { u = c[0] / c[3] * (0.5f * (float)p.W); v = c[1] / c[3] * (0.5f * (float)p.H); }
        }
        const unsigned bf = __ballot_sync(0xffffffffu, front);
        const bool allFront = (bf & gm) == gm;
        // (every ballot is taken by the whole warp before the results are combined: the 8-lane groups differ)
        const unsigned bxr = __ballot_sync(0xffffffffu, oxr), bxl = __ballot_sync(0xffffffffu, oxl);
        const unsigned byt = __ballot_sync(0xffffffffu, oyt), byb = __ballot_sync(0xffffffffu, oyb);
        const bool out = (bxr & gm) == gm || (bxl & gm) == gm || (byt & gm) == gm || (byb & gm) == gm;
        // screen footprint of the AABB (finite when every corner is in front): a batch with large triangles keeps its warp
        // busy several times longer than the median -- those go first
        float umin = u, umax = u, vmin = v, vmax = v;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            umin = fminf(umin, __shfl_xor_sync(0xffffffffu, umin, o)); umax = fmaxf(umax, __shfl_xor_sync(0xffffffffu, umax, o));
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o)); vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        }
        const bool visible = b < total && !(allFront && out) && !emptyBox;
        const bool leader = k == 0;
        const bool heavy = visible && (!allFront || (umax - umin) * (vmax - vmin) >= p.heavyArea);
        const unsigned bh = __ballot_sync(0xffffffffu, leader && heavy), bl_ = __ballot_sync(0xffffffffu, leader && visible && !heavy);
        __shared__ unsigned s_cnt[8][2], s_baseHL[2];
        const int warp = threadIdx.x >> 5;
        if (lane == 0) { s_cnt[warp][0] = __popc(bh); s_cnt[warp][1] = __popc(bl_); }
        __syncthreads();
        if (threadIdx.x < 2) {
            unsigned t = 0;
            for (int w = 0; w < 8; w++) t += s_cnt[w][threadIdx.x];
            s_baseHL[threadIdx.x] = t ? atomicAdd(threadIdx.x == 0 ? &p.ctr->nBatchHeavy : &p.ctr->nBatchLight, t) : 0u;
        }
        __syncthreads();
        if (leader && visible) {
            unsigned o = s_baseHL[heavy ? 0 : 1];
            for (int w = 0; w < warp; w++) o += s_cnt[w][heavy ? 0 : 1];
            o += __popc((heavy ? bh : bl_) & ((1u << lane) - 1u));
            p.batchList[heavy ? o : (unsigned)total - 1u - o] = (uint32_t)b;
        }
        if (threadIdx.x == 0) EHB_TL_STOP(p, 1, blockIdx.x, tl0);
        return;
    }
    blk -= batchBlocks;
    // ---------------------------------------------------------------- (C) clear
    // What the PREVIOUS pass on this scratch dirtied: planes [0, prevCursor) := EMPTY, bits [0, prevBitCursor) := 0 (k_raster
    // notes a pass's extents; beyond them the pools are still clean, a new allocation is cleared by the host).  Nothing here
    // depends on this pass's table, so the clear runs beside it.
    if (blk < clearBlocks) {
        const unsigned long long total = min(p.ctr->prevCursor, p.poolCap);
        const unsigned long long n2 = (total + 1ull) >> 1;
        ulonglong2* p2 = reinterpret_cast<ulonglong2*>(p.pool);
        for (unsigned long long i = (unsigned long long)blk * blockDim.x + threadIdx.x; i < n2;
             i += (unsigned long long)clearBlocks * blockDim.x)
            p2[i] = make_ulonglong2(EHB_EMPTY, EHB_EMPTY);
        if (p.bits) {
            const unsigned long long nb2 = (min(p.ctr->prevBitCursor, p.bitCap) + 1ull) >> 1;
            ulonglong2* b2 = reinterpret_cast<ulonglong2*>(p.bits);
            for (unsigned long long i = (unsigned long long)blk * blockDim.x + threadIdx.x; i < nb2;
                 i += (unsigned long long)clearBlocks * blockDim.x)
                b2[i] = make_ulonglong2(0ull, 0ull);
        }
        if (p.touch)
            for (int i = blk * blockDim.x + threadIdx.x; i < p.items * p.ntiles; i += clearBlocks * blockDim.x) p.touch[i] = 0u;
        if (threadIdx.x == 0) EHB_TL_STOP(p, 1, blockIdx.x, tl0);
        return;
    }
    blk -= clearBlocks;
    // ---------------------------------------------------------------- (L) tile lists
    // one LANE per tile; the three list counters get one atomic per warp each (a tile per warp meant ten thousand
    // same-address atomics per pass, which serialise on their L2 line)
    ehb_wait_table(p);
    const int wid = blk * blockDim.x + threadIdx.x;
    const bool valid = wid < p.items * p.ntiles;
    int nhit = 0;
    if (valid) {
        const int item = wid / p.ntiles, tile = wid - item * p.ntiles;
        const int tx = tile % p.ntx, ty = tile / p.ntx;
        const int rx0 = tx * EHB_T - p.hlo, ry0 = ty * EHB_T - p.hlo;
        const int rx1 = tx * EHB_T + EHB_T - 1 + p.hhi, ry1 = ty * EHB_T + EHB_T - 1 + p.hhi;
        for (int l = 0; l < p.Lp; l++) {
            const int4 pl = __ldcg(reinterpret_cast<const int4*>(&p.plane[(size_t)item * p.Lp + l]));   // x0, y0, w, h
            nhit += pl.z > 0 && pl.x <= rx1 && pl.x + pl.z - 1 >= rx0 && pl.y <= ry1 && pl.y + pl.w - 1 >= ry0;
        }
    }
    // tiles with several links take several times longer downstream: listed first (front), the rest from the back
    const bool heavy = valid && nhit >= 2, light = valid && nhit == 1, empty = valid && nhit == 0 && p.mode != EHB_MODE_AA_BWD;
    const unsigned bh = __ballot_sync(0xffffffffu, heavy), bl = __ballot_sync(0xffffffffu, light), be = __ballot_sync(0xffffffffu, empty);
    unsigned baseH = 0, baseL = 0, baseE = 0;
    if (lane == 0) {
        if (bh) baseH = atomicAdd(&p.ctr->nTiles, (unsigned)__popc(bh));
        if (bl) baseL = atomicAdd(&p.ctr->nLight, (unsigned)__popc(bl));
        if (be) baseE = atomicAdd(&p.ctr->nEmpty, (unsigned)__popc(be));
    }
    baseH = __shfl_sync(0xffffffffu, baseH, 0); baseL = __shfl_sync(0xffffffffu, baseL, 0); baseE = __shfl_sync(0xffffffffu, baseE, 0);
    const unsigned below = (1u << lane) - 1u;
    if (heavy) p.tileList[baseH + __popc(bh & below)] = (uint32_t)wid;
    else if (light) p.tileList[(unsigned)(p.items * p.ntiles) - 1u - (baseL + __popc(bl & below))] = (uint32_t)wid;
    else if (empty) p.emptyList[baseE + __popc(be & below)] = (uint32_t)wid;
    if (threadIdx.x == 0) EHB_TL_STOP(p, 1, blockIdx.x, tl0);
}

// ------------------------------------------------------------------------------------------------ k_raster
#ifndef EHB_RWARPS
#define EHB_RWARPS 8         // warps per raster CTA; every warp works alone on batches of 32 triangles
#endif
#ifndef EHB_RGROUPS
#define EHB_RGROUPS 2        // groups of 32 rows per handed-on unit
#endif
#define EHB_BLK_WORDS (32 * 32 + 64)
#ifndef EHB_RMIN_BLOCKS
#define EHB_RMIN_BLOCKS (1024 / (EHB_RWARPS * 32))
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

__device__ __forceinline__ int ehb_floor_to_int(float x) { return __float2int_rd(x); }
__device__ __forceinline__ int ehb_floor_to_int(double x) { return __double2int_rd(x); }

// Exact span of covered samples in one row of a triangle's clipped bbox.  Edge k along the row is
// C_k(dx) = R_k + ax_k * dx (threshold folded in: covered <=> C_k >= 0 for all k), monotone in dx, so the
// covered set is an interval [lo, hi] (empty when lo > hi).  With a = |ax|:  ax < 0 bounds dx <= floor(R / a),
// ax > 0 bounds dx >= ceil(-R / a) = -floor(R / a), ax == 0 is all or nothing.  floor(R / a) is estimated with one
// float division (within 0.004 of the quotient for |quotient| <= w + 1 <= 8193; clamped beyond, where either answer
// lies outside the row) and fixed up with one exact integer remainder, so the result equals the brute-force test of
// every sample.  No branch depends on the edge's orientation: the lanes of a warp hold rows of different triangles.
template <typename I, typename F>
__device__ __forceinline__ void ehb_row_span(const I R0, const I R1, const I R2, const I ax0, const I ax1, const I ax2,
                                             int w, int& lo, int& hi)
{
    lo = 0; hi = w - 1;
    const I R[3] = {R0, R1, R2}, ax[3] = {ax0, ax1, ax2};
    const F lim = (F)(w + 1);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const I a = ax[k] < 0 ? -ax[k] : ax[k];
        const I a1 = a > 0 ? a : (I)1;
        F q = ehb_fast_div((F)R[k], (F)a1);
        q = fmin(fmax(q, -lim), lim);
        int fl = ehb_floor_to_int(q);
        const I rem = R[k] - a1 * (I)fl;
        fl += (rem >= a1 ? 1 : 0) - (rem < 0 ? 1 : 0);
        const int hik = ax[k] < 0 ? fl : ((ax[k] == 0 && R[k] < 0) ? -1 : w - 1);
        const int lok = ax[k] > 0 ? -fl : 0;
        hi = min(hi, hik); lo = max(lo, lok);
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

// Setup of one triangle given by its snapped vertices (1/16 px) -> record.  c0..c2 = the ORIGINAL clip-space vertices
// (the depth of a sample is shaded from them, also for the sub-triangles of a clipped triangle).  Same tests, in the same
// order, as the oracle's raster_snapped.  Returns the number of rows of the clipped bbox (0: nothing to draw).
__device__ __forceinline__ int ehb_setup_record(const EhbParams& p, const EhbPlane& pl, int x0, int y0, int x1, int y1, int x2, int y2,
                                                const float4& c0, const float4& c1, const float4& c2, uint32_t id, EhbRec& rc)
{
    const long long area = (long long)(x1 - x0) * (y2 - y0) - (long long)(y1 - y0) * (x2 - x0);
    if (area == 0) return 0;
    if (area < 0) { int t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
    const int bx = 8 * p.W - 8, by = 8 * p.H - 8;
    const int pxlo = max((min(x0, min(x1, x2)) + bx + 15) >> 4, 0), pxhi = min((max(x0, max(x1, x2)) + bx) >> 4, p.W - 1);
    const int pylo = max((min(y0, min(y1, y2)) + by + 15) >> 4, 0), pyhi = min((max(y0, max(y1, y2)) + by) >> 4, p.H - 1);
    if (pxlo > pxhi || pylo > pyhi) return 0;
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
    rc.id = id;
    rc.clip[0] = c0.x; rc.clip[1] = c0.y; rc.clip[2] = c0.z; rc.clip[3] = c0.w;
    rc.clip[4] = c1.x; rc.clip[5] = c1.y; rc.clip[6] = c1.z; rc.clip[7] = c1.w;
    rc.clip[8] = c2.x; rc.clip[9] = c2.y; rc.clip[10] = c2.z; rc.clip[11] = c2.w;
    return rc.h;
}

// Primitive assembly of one triangle from the pre-transformed vertices -> record.  Same tests, in the same order, as the
// oracle's eho_rasterize.  Returns the number of rows of its clipped bbox (0: nothing to draw here); needClip: the
// triangle survives the trivial rejection but leaves the depth range or the fixed-point guard band -- it goes through the
// clipper (ehb_emit_clipped).
__device__ __forceinline__ int ehb_make_record(const EhbRobot& rb, const EhbParams& p, int item, int l, int f, EhbRec& rc, bool& needClip)
{
    needClip = false;
    const int g = rb.foff[l] + f;
    const EhbLink& lk = rb.link[l];
    const int4 id = __ldg(lk.faces + f);
    const EhbPlane pl = p.plane[(size_t)item * p.Lp + (p.Lp == 1 ? 0 : l)];   // depends on the link only: issued with the face
    if ((unsigned)id.x >= (unsigned)lk.V || (unsigned)id.y >= (unsigned)lk.V || (unsigned)id.z >= (unsigned)lk.V) return 0;
    const size_t vb = (size_t)item * p.Vtot + rb.voff[l];
    // all seven loads that depend on the face go out together (one L2 round trip), before the first test consumes one
    const float4 c0 = p.vclip[vb + id.x], c1 = p.vclip[vb + id.y], c2 = p.vclip[vb + id.z];
    const int2 s0 = p.vsnap[vb + id.x], s1 = p.vsnap[vb + id.y], s2 = p.vsnap[vb + id.z];
    if ((c0.w < c0.x && c1.w < c1.x && c2.w < c2.x) || (c0.w < -c0.x && c1.w < -c1.x && c2.w < -c2.x) ||
        (c0.w < c0.y && c1.w < c1.y && c2.w < c2.y) || (c0.w < -c0.y && c1.w < -c1.y && c2.w < -c2.y) ||
        (c0.w < c0.z && c1.w < c1.z && c2.w < c2.z) || (c0.w < -c0.z && c1.w < -c1.z && c2.w < -c2.z))
        return 0;
    const int G = 1 << 28;
    if (s0.x == INT_MIN || s1.x == INT_MIN || s2.x == INT_MIN ||   // a vertex outside the depth range
        s0.x > G || s0.x < -G || s0.y > G || s0.y < -G || s1.x > G || s1.x < -G || s1.y > G || s1.y < -G ||
        s2.x > G || s2.x < -G || s2.y > G || s2.y < -G) {
        needClip = true;
        return 0;
    }
    return ehb_setup_record(p, pl, s0.x, s0.y, s1.x, s1.y, s2.x, s2.y, c0, c1, c2, p.Lp == 1 ? (uint32_t)g : (uint32_t)f, rc);
}

// Sutherland-Hodgman clip of a triangle against the six planes of the view frustum in clip space (x >= -w, x <= w,
// y >= -w, y <= w, z >= -w, z <= w), fp32, one rounding per operation; an intersection is always computed from the
// INSIDE vertex towards the outside one, so that neighbours are cut at bit-identical points (oracle: clip_triangle).
__device__ __forceinline__ float ehb_clip_dist(const float* v, int plane)
{
    const float c = plane < 2 ? v[0] : (plane < 4 ? v[1] : v[2]);
    return (plane & 1) ? v[3] - c : v[3] + c;
}
__device__ __noinline__ int ehb_clip_triangle(const float4 c0, const float4 c1, const float4 c2, float (*out)[4])
{
    float a[9][4], b[9][4];
    int n = 3;
    a[0][0] = c0.x; a[0][1] = c0.y; a[0][2] = c0.z; a[0][3] = c0.w;
    a[1][0] = c1.x; a[1][1] = c1.y; a[1][2] = c1.z; a[1][3] = c1.w;
    a[2][0] = c2.x; a[2][1] = c2.y; a[2][2] = c2.z; a[2][3] = c2.w;
    for (int plane = 0; plane < 6 && n >= 3; plane++) {
        int m = 0;
        for (int i = 0; i < n; i++) {
            const float* pp = a[i];
            const float* qq = a[(i + 1) % n];
            const float dp = ehb_clip_dist(pp, plane), dq = ehb_clip_dist(qq, plane);
            const bool ip = dp >= 0.f, iq = dq >= 0.f;
            if (ip && m < 9) { for (int k = 0; k < 4; k++) b[m][k] = pp[k]; m++; }
            if (ip != iq && m < 9) {
                const float* in = ip ? pp : qq;
                const float* ou = ip ? qq : pp;
                const float din = ip ? dp : dq, dou = ip ? dq : dp;
                const float t = din / (din - dou);
                for (int k = 0; k < 4; k++) b[m][k] = in[k] + t * (ou[k] - in[k]);
                m++;
            }
        }
        n = m;
        for (int i = 0; i < n; i++)
            for (int k = 0; k < 4; k++) a[i][k] = b[i][k];
    }
    if (n < 3) return 0;
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 4; k++) out[i][k] = a[i][k];
    return n;
}

// One lane draws one record by testing every sample of its bbox: only when the deferred-work queues are full.
__device__ __noinline__ void ehb_draw_serial(const EhbParams& p, const EhbRec& rc, const EhbBits bits)
{
    const float p0[4] = {rc.clip[0], rc.clip[1], rc.clip[2], rc.clip[3]}, p1[4] = {rc.clip[4], rc.clip[5], rc.clip[6], rc.clip[7]},
                p2[4] = {rc.clip[8], rc.clip[9], rc.clip[10], rc.clip[11]};
    for (int dy = 0; dy < rc.h; dy++)
        for (int dx = 0; dx < rc.w; dx++) {
            bool in = true;
            for (int k = 0; k < 3; k++)
                in = in && (rc.E[k] + 16ll * rc.ex[k] * (long long)dy - 16ll * rc.ey[k] * (long long)dx) >= 0;
            if (!in) continue;
            const int px = rc.x0 + dx, py = rc.y0 + dy;
            const float zw = ehb_shade_zw(p0, p1, p2, p.xs * (float)px + p.xo, p.ys * (float)py + p.yo);
            atomicMin(p.pool + (rc.base + (long long)py * rc.pw + px), ((unsigned long long)ehb_order_key(zw) << 32) | rc.id);
            if (bits.w) ehb_bits_set(bits, px, py, 1);
        }
}

// A triangle that needs the clipper (one lane, rare): clipped polygon -> fan of sub-triangles -> each one a deferred
// record whose bbox is cut into bounded units for k_raster_big (a triangle that crosses the near plane can cover the
// whole screen), with the link's touch bits of the tiles it reaches.
__device__ __noinline__ void ehb_emit_clipped(const EhbRobot& rb, const EhbParams& p, int item, int l, int f, int qi)
{
    atomicAdd(&p.ctr->nNeedClip, 1ull);
    ehb_raise(p, 2u);
    const EhbLink& lk = rb.link[l];
    const int4 id = __ldg(lk.faces + f);
    const EhbPlane pl = p.plane[(size_t)item * p.Lp + (p.Lp == 1 ? 0 : l)];
    const EhbBits bits = ehb_bits_of(p.bits, pl);
    const size_t vb = (size_t)item * p.Vtot + rb.voff[l];
    const float4 c0 = p.vclip[vb + id.x], c1 = p.vclip[vb + id.y], c2 = p.vclip[vb + id.z];
    float poly[9][4];
    const int n = ehb_clip_triangle(c0, c1, c2, poly);
    int sx[9], sy[9];
    for (int i = 0; i < n; i++) {
        if (!(poly[i][3] > 0.f)) return;
        const float r = 1.0f / poly[i][3];
        sx[i] = ehb_rni_sat(poly[i][0] * r * (float)(p.W * 8));
        sy[i] = ehb_rni_sat(poly[i][1] * r * (float)(p.H * 8));
    }
    EhbCounters::Q& myq = p.ctr->q[qi];
    const uint32_t tid = p.Lp == 1 ? (uint32_t)(rb.foff[l] + f) : (uint32_t)f;
    for (int i = 1; i + 1 < n; i++) {
        EhbRec rc;
        if (ehb_setup_record(p, pl, sx[0], sy[0], sx[i], sy[i], sx[i + 1], sy[i + 1], c0, c1, c2, tid, rc) == 0) continue;
        if (p.touch) {
            const int txlo = max(0, (rc.x0 - p.hhi) >> 5), txhi = min(p.ntx - 1, (rc.x0 + rc.w - 1 + p.hlo) >> 5);
            const int tylo = max(0, (rc.y0 - p.hhi) >> 5), tyhi = min(p.nty - 1, (rc.y0 + rc.h - 1 + p.hlo) >> 5);
            uint32_t* trow = p.touch + (size_t)item * p.ntiles;
            for (int ty = tylo; ty <= tyhi; ty++)
                for (int tx = txlo; tx <= txhi; tx++) atomicOr(trow + ty * p.ntx + tx, 1u << l);
        }
        const int nux = (rc.w + EHB_UNIT_W - 1) / EHB_UNIT_W, nuy = (rc.h + EHB_UNIT_H - 1) / EHB_UNIT_H, nu = nux * nuy;
        const unsigned kq = atomicAdd(&myq.nBigRec, 1u), uq = atomicAdd(&myq.nUnits, (unsigned)nu);
        const bool fits = (int)kq < p.bigCap && (int)(uq + nu) <= p.unitCap;
        const unsigned u0 = (unsigned)qi * (unsigned)p.unitCap + uq;
        if (fits) {
            const unsigned k = (unsigned)qi * (unsigned)p.bigCap + kq;
            p.bigRec[k] = rc;
            if (p.bigBits) p.bigBits[k] = EhbBitsRec{bits.base, bits.pitch, 0};
            for (int uy = 0; uy < nuy; uy++)
                for (int ux = 0; ux < nux; ux++)
                    p.units[u0 + uy * nux + ux] = EhbUnit{k, (unsigned short)(ux * EHB_UNIT_W), (unsigned short)(uy * EHB_UNIT_H)};
        } else {
            ehb_raise(p, 4u);   // queues full: drawn here (slow but complete); its units are void
            for (int u = 0; u < nu && (int)(uq + u) < p.unitCap; u++) p.units[u0 + u] = EhbUnit{0xFFFFFFFFu, 0, 0};
            ehb_draw_serial(p, rc, bits);
        }
    }
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
// `t` = this lane's record index in recs (-1: no row), `dy` = its row inside that record's bbox.  `cmp` = 32 words-pairs of
// the warp's shared memory.  The covered runs of the rows form one flat list of samples; sample j belongs to the last
// non-empty row that starts at or before j.  The non-empty rows are compacted into `cmp` (first sample, start offset) once
// per group; per 32 samples ONE warp-wide OR (REDUX) of "my row starts at sample s of this chunk" bits and a popcount give
// every lane its row -- instead of a five-step binary search over the row prefix by shuffles (a third of the loop's
// instructions).  SINGLE: every row belongs to record 0 (a deferred triangle's window): its fields stay in registers.
template <typename I, typename F, class RV, bool SINGLE>
__device__ __forceinline__ void ehb_rows_group(const RV rv, int t, int dy, int lane, unsigned long long* pool,
                                               float xs, float xo, float ys, float yo, const EhbBits bits, uint2* cmp,
                                               int cx0 = 0, int cx1 = 1 << 20)
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
        // the covered run of this row -> the plane's coverage bits (one or two RED.OR); the depth samples follow below
        if (bits.w && len > 0) ehb_bits_set(bits, rv.i(t, 12) + a, rv.i(t, 13) + dy, len);
    }
    const unsigned ne = __ballot_sync(0xffffffffu, len > 0);
    if (ne == 0u) return;                                  // nothing covered in these rows
    int inc = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    const int exc = inc - len;
    __syncwarp();                                          // (the previous group's readers are done with cmp)
    if (len > 0) cmp[__popc(ne & ((1u << lane) - 1u))] = make_uint2(pos, (uint32_t)exc);
    __syncwarp();
    // SINGLE: the one record's fields, loaded once
    float s0[4], s1[4], s2[4];
    long long sbase = 0; int spw = 0; uint32_t sid = 0;
    if (SINGLE) {
#pragma unroll
        for (int k = 0; k < 4; k++) { s0[k] = rv.f(0, 20 + k); s1[k] = rv.f(0, 24 + k); s2[k] = rv.f(0, 28 + k); }
        sbase = rv.ll(0, 16); spw = rv.i(0, 18); sid = rv.u(0, 19);
    }
    int kbase = 0;                                         // non-empty rows that start before the chunk
    const unsigned upto = 0xFFFFFFFFu >> (31 - lane);      // bits 0 .. lane
    for (int j0 = 0; j0 < total; j0 += 32) {
        const unsigned rel = (unsigned)(exc - j0);
        const unsigned heads = __reduce_or_sync(0xffffffffu, (len > 0 && rel < 32u) ? (1u << rel) : 0u);
        const int k = kbase + __popc(heads & upto) - 1;
        kbase += __popc(heads);
        const int j = j0 + lane;
        if (j < total) {
            const uint2 c = cmp[k];
            const uint32_t q = c.x + (uint32_t)(j - (int)c.y);
            const int px = (int)(q & 8191u), py = (int)((q >> 13) & 8191u);
            if (SINGLE) {
                const float zw = ehb_shade_zw(s0, s1, s2, xs * (float)px + xo, ys * (float)py + yo);
                atomicMin(pool + (sbase + (long long)py * spw + px), ((unsigned long long)ehb_order_key(zw) << 32) | sid);
            } else {
                ehb_shade_global(rv, (int)(q >> 26), px, py, pool, xs, xo, ys, yo);
            }
        }
    }
}

__global__ void __launch_bounds__(EHB_RWARPS * 32, EHB_RMIN_BLOCKS) ehb_k_raster(const __grid_constant__ EhbRobot rb,
                                                                const __grid_constant__ EhbParams p, int streamBlocks,
                                                                int chunks)
{
    ehb_pdl_enter();
    EHB_TL_START(tl0);
    __shared__ __align__(128) uint32_t s_rec[EHB_RWARPS][32 * 32];   // 32 records per warp, transposed
    __shared__ int s_off[EHB_RWARPS][33];
    __shared__ uint2 s_cmp[EHB_RWARPS][32];                           // compacted non-empty rows of a group (ehb_rows_group)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if ((int)blockIdx.x < streamBlocks) {
        // spare CTAs of this launch finish the tiles no link touches: pure HBM streaming that overlaps the
        // instruction-bound rasterization instead of sitting in front of it
        const int n = (int)p.ctr->nEmpty;
        for (int i = blockIdx.x * EHB_RWARPS + warp; i < n; i += streamBlocks * EHB_RWARPS) {
            const int wid = (int)p.emptyList[i];
            const int item = wid / p.ntiles, tile = wid - item * p.ntiles;
            ehb_stream_empty_tile(p, item, tile % p.ntx, tile / p.ntx, lane);
        }
        return;
    }
    // One warp per listed batch (32 consecutive faces of one link of one item): k_front listed the batches that survive the
    // frustum test of their AABB, the ones with a large screen footprint first.  The grid covers every batch of the pass;
    // the warps beyond the list leave at once.
    if (blockIdx.x == (unsigned)streamBlocks && threadIdx.x == 0) {
        p.ctr->tableReady = 0u;                          // k_front is complete: re-arm its flag,
        p.ctr->prevCursor = p.ctr->planeCursor;          // note what this pass dirties (the next k_front clears it)
        if (p.bits) p.ctr->prevBitCursor = p.ctr->bitCursor;
    }
    EHB_MARK(p, 8);
    const int total = chunks * p.items;   // chunks = 32-triangle batches per item
    const unsigned e = (unsigned)(((int)blockIdx.x - streamBlocks) * EHB_RWARPS + warp);
    const unsigned listed = e < (unsigned)total ? p.batchList[e] : 0u;               // (speculative: issued with the counters)
    const unsigned nHeavy = p.ctr->nBatchHeavy, nLight = p.ctr->nBatchLight;
    if (e >= nHeavy + nLight) return;
    const unsigned bcur = e < nHeavy ? listed : p.batchList[(unsigned)total - 1u - (e - nHeavy)];
    const float xs = p.xs, xo = p.xo, ys = p.ys, yo = p.yo;
    const EhbRecSoA recs{s_rec[warp]};
    int* off = s_off[warp];
    uint2* cmp = s_cmp[warp];
    {
    const int qi = (int)(bcur & (EHB_NQ - 1));            // this batch's sub-queue
    EhbCounters::Q& myq = p.ctr->q[qi];
    const int item = (int)bcur / chunks;
    const int bl = (int)bcur - item * chunks;             // batch of the item
    const int link = ehb_find_link(rb.boff, rb.L, bl);    // (warp-uniform)
    const int f = (bl - rb.boff[link]) * 32 + lane;
    int rows = 0, wide = 0, touchRows = 0;
    bool big = false, needClip = false;
    int4 tb = make_int4(0, 0, 0, 0);   // clipped bbox of this lane's triangle (x0, y0, w, h)
    const bool visible = true;
    // the coverage bits of the batch's plane (one link of one item: the same for every lane)
    const EhbBits bits = ehb_bits_of(p.bits, p.plane[(size_t)item * p.Lp + (p.Lp == 1 ? 0 : link)]);
    if (visible && f < rb.link[link].F) {
        EhbRec rc;
        rows = ehb_make_record(rb, p, item, link, f, rc, needClip);
        touchRows = rows;
        if (rows > 0) tb = make_int4(rc.x0, rc.y0, rc.w, rc.h);
        if (rows > 0) {
            const int ext = max(max(abs(rc.ex[0]), abs(rc.ex[1])), max(max(abs(rc.ex[2]), abs(rc.ey[0])), max(abs(rc.ey[1]), abs(rc.ey[2]))));
            wide = ext >= 32768;   // 32-bit edge arithmetic is exact below 2^15 sub-pixel units per edge
            big = rc.w * rc.h > p.smallArea;   // not small: park the record, cut the bbox into bounded units
            ehb_rec_store_soa(s_rec[warp], lane, rc);
        }
    }
    {   // deferred triangles of this batch: one pair of queue atomics per warp, records copied out by the whole warp
        const unsigned bm = __ballot_sync(0xffffffffu, big);
        if (bm) {
            const int nux = big ? (tb.z + EHB_UNIT_W - 1) / EHB_UNIT_W : 0, nuy = big ? (tb.w + EHB_UNIT_H - 1) / EHB_UNIT_H : 0;
            const int nu = nux * nuy;
            int uinc = nu;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, uinc, o);
                if (lane >= o) uinc += v;
            }
            const int utot = __shfl_sync(0xffffffffu, uinc, 31);
            unsigned k0 = 0, ub = 0;
            if (lane == 0) { k0 = atomicAdd(&myq.nBigRec, (unsigned)__popc(bm)); ub = atomicAdd(&myq.nUnits, (unsigned)utot); }
            k0 = __shfl_sync(0xffffffffu, k0, 0); ub = __shfl_sync(0xffffffffu, ub, 0);
            const unsigned kq = k0 + (unsigned)__popc(bm & ((1u << lane) - 1u));   // index in the sub-queue
            const unsigned uq = ub + (unsigned)(uinc - nu);
            const bool fits = (int)kq < p.bigCap && (int)(uq + nu) <= p.unitCap;
            const unsigned k = (unsigned)qi * (unsigned)p.bigCap + kq;             // index in the arrays
            const unsigned u0 = (unsigned)qi * (unsigned)p.unitCap + uq;
            if (big) {
                if (fits) {
                    // A window of the bbox that lies entirely outside one edge (the edge function's maximum over the
                    // window is negative) holds no covered sample: it becomes a void unit.  Sliver triangles have many.
                    long long E[3]; int ex[3], ey[3];
#pragma unroll
                    for (int q = 0; q < 3; q++) { E[q] = recs.ll(lane, 2 * q); ex[q] = recs.i(lane, 6 + q); ey[q] = recs.i(lane, 9 + q); }
                    for (int uy = 0; uy < nuy; uy++)
                        for (int ux = 0; ux < nux; ux++) {
                            const int dx0 = ux * EHB_UNIT_W, dx1 = min(tb.z - 1, dx0 + EHB_UNIT_W - 1);
                            const int dy0 = uy * EHB_UNIT_H, dy1 = min(tb.w - 1, dy0 + EHB_UNIT_H - 1);
                            bool any = true;
#pragma unroll
                            for (int q = 0; q < 3; q++) {
                                const long long mx = E[q] + 16ll * ex[q] * (long long)(ex[q] > 0 ? dy1 : dy0) -
                                                     16ll * ey[q] * (long long)(ey[q] > 0 ? dx0 : dx1);
                                any = any && mx >= 0;
                            }
                            p.units[u0 + uy * nux + ux] = any ? EhbUnit{k, (unsigned short)dx0, (unsigned short)dy0} : EhbUnit{0xFFFFFFFFu, 0, 0};
                        }
                    rows = 0;
                } else {
                    ehb_raise(p, 4u);   // queues full: this one is drawn inline (slow but complete); its units are void
                    for (int i = 0; i < nu && (int)(uq + i) < p.unitCap; i++) p.units[u0 + i] = EhbUnit{0xFFFFFFFFu, 0, 0};
                }
            }
            __syncwarp();
            // park the records: lane w writes word w of each deferred record (one 128-B store per record)
            unsigned todo = __ballot_sync(0xffffffffu, big && fits);
            while (todo) {
                const int t = __ffs(todo) - 1;
                todo &= todo - 1;
                const unsigned kt = __shfl_sync(0xffffffffu, k, t);
                reinterpret_cast<uint32_t*>(p.bigRec + kt)[lane] = s_rec[warp][lane * 32 + t];
                if (lane == 0 && p.bigBits) p.bigBits[kt] = EhbBitsRec{bits.base, bits.pitch, 0};
            }
        }
    }
    if (p.touch) {
        // Tell k_tiles which (tile, link) windows the triangles reach into.  Neighbouring triangles of a warp mostly
        // fall into the same tile of the same link: lanes with equal (tile, link) elect one to issue the RED (no load,
        // nothing waits for it); bboxes that straddle a tile border (halo included) set every tile they reach.
        int txlo = 0, txhi = -1, tylo = 0, tyhi = -1;
        if (touchRows > 0) {
            txlo = max(0, (tb.x - p.hhi) >> 5); txhi = min(p.ntx - 1, (tb.x + tb.z - 1 + p.hlo) >> 5);
            tylo = max(0, (tb.y - p.hhi) >> 5); tyhi = min(p.nty - 1, (tb.y + tb.w - 1 + p.hlo) >> 5);
        }
        const bool single = touchRows > 0 && txlo == txhi && tylo == tyhi;
        const uint32_t key = single ? (((uint32_t)(tylo * p.ntx + txlo) << 5) | (uint32_t)link) : (0xFFFFFFE0u | (uint32_t)lane);
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        uint32_t* trow = p.touch + (size_t)item * p.ntiles;
        if (single) {
            if ((int)(__ffs(peers) - 1) == lane) atomicOr(trow + tylo * p.ntx + txlo, 1u << link);
        } else {
            for (int ty = tylo; ty <= tyhi; ty++)
                for (int tx = txlo; tx <= txhi; tx++) atomicOr(trow + ty * p.ntx + tx, 1u << link);
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
    const int nRowsAll = off[32];
    // A heavy batch would make this warp the tail of the launch: it keeps p.inlineGroups groups of 32 rows and hands the
    // rest to k_raster_big, which spreads such units over the whole chip (records + row prefix parked in global).  The
    // queue tickets are drawn before the inline groups and used after them, so nothing waits for the atomics.
    const bool heavy = nRowsAll > 32 * p.inlineGroups && p.batchBlk != nullptr;
    const int ngroups = (nRowsAll + 31) >> 5;
    const int nitems = heavy ? (ngroups - p.inlineGroups + EHB_RGROUPS - 1) / EHB_RGROUPS : 0;
    unsigned slot = 0, u0 = 0;
    if (heavy && lane == 0) { slot = atomicAdd(&myq.nBatchBlk, 1u); u0 = atomicAdd(&myq.nUnits, (unsigned)nitems); }
    const int nInline = heavy ? 32 * p.inlineGroups : nRowsAll;
    auto draw_rows = [&](int rBegin, int rEnd) {
        for (int r0 = rBegin; r0 < rEnd; r0 += 32) {
            const int r = r0 + lane;
            int t = -1, dy = 0;
            if (r < rEnd) {
                int lo = 0, hi = 32;   // last t with off[t] <= r
#pragma unroll
                for (int st = 0; st < 5; st++) {
                    const int mid = (lo + hi) >> 1;
                    if (off[mid] <= r) lo = mid; else hi = mid;
                }
                t = lo; dy = r - off[t];
            }
            if (anyWide) ehb_rows_group<long long, double, EhbRecSoA, false>(recs, t, dy, lane, p.pool, xs, xo, ys, yo, bits, cmp);
            else ehb_rows_group<int, float, EhbRecSoA, false>(recs, t, dy, lane, p.pool, xs, xo, ys, yo, bits, cmp);
        }
    };
    draw_rows(0, nInline);
    if (heavy) {
        slot = __shfl_sync(0xffffffffu, slot, 0); u0 = __shfl_sync(0xffffffffu, u0, 0);
        const bool fits = (int)slot < p.batchCap && (int)(u0 + nitems) <= p.unitCap;
        const unsigned uBase = (unsigned)qi * (unsigned)p.unitCap + u0;
        slot += (unsigned)qi * (unsigned)p.batchCap;
        if (fits) {
            uint32_t* blk = p.batchBlk + (size_t)slot * EHB_BLK_WORDS;
#pragma unroll 8
            for (int k = 0; k < 32; k++) blk[k * 32 + lane] = s_rec[warp][k * 32 + lane];
            blk[1024 + lane] = (uint32_t)off[lane];
            if (lane == 0) {
                blk[1024 + 32] = (uint32_t)nRowsAll; blk[1024 + 33] = anyWide ? 1u : 0u;
                blk[1024 + 34] = (uint32_t)(unsigned long long)bits.base; blk[1024 + 35] = (uint32_t)((unsigned long long)bits.base >> 32);
                blk[1024 + 36] = (uint32_t)bits.pitch;
            }
            for (int i = lane; i < nitems; i += 32)
                p.units[uBase + i] = EhbUnit{0x80000000u | slot, (unsigned short)(p.inlineGroups + i * EHB_RGROUPS), (unsigned short)EHB_RGROUPS};
        } else {   // no room: the units are void and the rest of the batch is drawn here
            for (int i = lane; i < nitems; i += 32)
                if ((int)(u0 + i) < p.unitCap) p.units[uBase + i] = EhbUnit{0xFFFFFFFFu, 0, 0};
            draw_rows(nInline, nRowsAll);
        }
    }
    // (rare) triangles of this batch that left the depth range / guard band: one lane each clips, sets up and queues the
    // sub-triangles -- at the end of the iteration, where nothing of the batch is live any more
    if (__any_sync(0xffffffffu, needClip)) {
        if (needClip) ehb_emit_clipped(rb, p, item, link, f, qi);
    }
    if (lane == 0) EHB_TL_STOP(p, 2, ((int)blockIdx.x - streamBlocks) * EHB_RWARPS + warp, tl0);
    }
}

// Deferred work: the warps draw units from the 32 sub-queues; one unit = a 64 x 32 pixel window of one triangle's bbox, or
// two groups of 32 rows of a parked batch.

#ifndef EHB_BMIN_BLOCKS
#define EHB_BMIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(256, EHB_BMIN_BLOCKS) ehb_k_raster_big(const __grid_constant__ EhbParams p)
{
    ehb_pdl_enter();
    EHB_TL_START(tl0);
    __shared__ __align__(128) uint32_t s_blk[8][EHB_BLK_WORDS];
    __shared__ __align__(8) uint64_t s_bar[8];           // one mbarrier per warp: completion of its bulk copies
    __shared__ uint2 s_cmp[8][32];                       // compacted non-empty rows of a group (ehb_rows_group)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0) { p.ctr->nBatchHeavy = 0u; p.ctr->nBatchLight = 0u; }   // k_raster is complete
    EHB_MARK(p, 9);
    if (lane == 0) ehb_mbar_init(&s_bar[warp], 1);
    ehb_fence_mbar_init();
    __syncwarp();
    uint32_t phase = 0;
    const int qn = min((int)p.ctr->q[lane].nUnits, p.unitCap);   // lane s: units of sub-queue s
    const float xs = p.xs, xo = p.xo, ys = p.ys, yo = p.yo;
    uint32_t* sb = s_blk[warp];
    uint2* cmp = s_cmp[warp];
    // a block of a warp's shared memory <- global memory as ONE bulk asynchronous copy (UBLKCP) that completes on the
    // warp's mbarrier: no register staging, one instruction instead of a load + store per word
    auto fetch_issue = [&](const void* src, uint32_t bytes) {
        __syncwarp();                          // every lane is done reading the previous contents
        if (lane == 0) {
            ehb_fence_proxy_async();           // ... and those reads are ordered before the asynchronous write
            ehb_mbar_arrive_expect_tx(&s_bar[warp], bytes);
            ehb_bulk_g2s(sb, src, bytes, &s_bar[warp]);
        }
    };
    auto fetch_wait = [&]() {
        ehb_mbar_wait(&s_bar[warp], phase);
        phase ^= 1u;
    };
    // Units are drawn dynamically: every sub-queue has a cursor (on the sub-queue's own L2 line) and 1/32 of the warps;
    // a warp takes the units of its sub-queue one at a time until they are used up -- the units' cost varies from nothing
    // (a sliver's window) to 2048 samples: with a static split a tenth of the warps were still drawing in the last 8 us.
    // The sub-queues themselves hold equal shares of the work (batches are dealt round-robin, ~500 units each), so nothing
    // is stolen across them: a look at the other cursors costs more than the imbalance (measured).  The ticket of the
    // next unit and its descriptor are fetched while the current unit is drawn.
    const int sq = (blockIdx.x * 8 + warp) & (EHB_NQ - 1);
    const int qs = __shfl_sync(0xffffffffu, qn, sq);
    const EhbUnit UNIT_END = EhbUnit{0xFFFFFFFEu, 0, 0};
    auto take = [&]() -> unsigned { return lane == 0 ? atomicAdd(&p.ctr->q[sq].take, 1u) : 0u; };
    auto resolve = [&](unsigned tk) -> EhbUnit {          // ticket (lane 0) -> unit descriptor
        const unsigned b = __shfl_sync(0xffffffffu, tk, 0);
        return (int)b < qs ? p.units[(size_t)sq * p.unitCap + b] : UNIT_END;
    };
    EhbUnit un = resolve(take());
    unsigned tk = take();
    for (;;) {
        const EhbUnit cur = un;
        if (cur.rec == UNIT_END.rec) break;
        // this unit's record(s) -> shared memory (bulk copy), and behind it -- after the copy's fence, which waits for
        // everything this lane has in flight -- the next unit's descriptor load and the ticket of the one after, so that both
        // travel while this unit is drawn (a ticket drawn past the end is harmless)
        const bool isVoid = cur.rec == 0xFFFFFFFFu, isBatch = !isVoid && (cur.rec & 0x80000000u);
        if (isBatch) fetch_issue(p.batchBlk + (size_t)(cur.rec & 0x7FFFFFFFu) * EHB_BLK_WORDS, EHB_BLK_WORDS * 4);
        else if (!isVoid) fetch_issue(p.bigRec + cur.rec, (uint32_t)sizeof(EhbRec));
        EhbBits ubits;
        ubits.w = nullptr; ubits.base = 0; ubits.pitch = 0;
        if (!isVoid && !isBatch && p.bigBits) {
            const EhbBitsRec br = p.bigBits[cur.rec];
            ubits.w = p.bits; ubits.base = br.base; ubits.pitch = br.pitch;
        }
        un = resolve(tk);
        tk = take();
        if (isVoid) continue;
        fetch_wait();
        if (isBatch) {
            // rows [32 * dx0, 32 * (dx0 + dy0)) of a parked batch: the same row-group loop as k_raster
            const int* off = reinterpret_cast<const int*>(sb + 1024);
            const int nRows = off[32];
            const bool anyWide = off[33] != 0;
            EhbBits bits;
            bits.w = p.bits; bits.base = (long long)(((unsigned long long)sb[1024 + 35] << 32) | sb[1024 + 34]); bits.pitch = off[36];
            const EhbRecSoA recs{sb};
            for (int gi = 0; gi < (int)cur.dy0; gi++) {
                const int r0 = ((int)cur.dx0 + gi) * 32;
                if (r0 >= nRows) break;
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
                if (anyWide) ehb_rows_group<long long, double, EhbRecSoA, false>(recs, t, dy, lane, p.pool, xs, xo, ys, yo, bits, cmp);
                else ehb_rows_group<int, float, EhbRecSoA, false>(recs, t, dy, lane, p.pool, xs, xo, ys, yo, bits, cmp);
            }
            continue;
        }
        // a 64 x 32 window of one deferred triangle
        const EhbRec* rc = reinterpret_cast<const EhbRec*>(sb);
        const int ext = max(max(abs(rc->ex[0]), abs(rc->ex[1])), max(max(abs(rc->ex[2]), abs(rc->ey[0])), max(abs(rc->ey[1]), abs(rc->ey[2]))));
        const int dy = cur.dy0 + lane;
        const int t = dy < rc->h ? 0 : -1;
        const EhbRecAoS rv{reinterpret_cast<const uint32_t*>(rc)};
        if (ext >= 32768) ehb_rows_group<long long, double, EhbRecAoS, true>(rv, t, dy, lane, p.pool, xs, xo, ys, yo, ubits, cmp, cur.dx0, cur.dx0 + EHB_UNIT_W - 1);
        else ehb_rows_group<int, float, EhbRecAoS, true>(rv, t, dy, lane, p.pool, xs, xo, ys, yo, ubits, cmp, cur.dx0, cur.dx0 + EHB_UNIT_W - 1);
    }
    if (lane == 0) EHB_TL_STOP(p, 3, blockIdx.x * 8 + warp, tl0);
    EHB_MARK(p, 12);
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
