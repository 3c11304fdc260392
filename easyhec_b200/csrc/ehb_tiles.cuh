// ehb_tiles.cuh -- the image-space half of a pass: antialias, compose, loss, backward -- ONE kernel.
//
// The reference antialiases every link on its own, sums the per-link masks and clamps (rb_solver.py:62-68), and its
// backward scatters one gradient per silhouette pixel pair (dr.antialias).  Here one small CTA (4 warps)
// owns one listed 32x32 tile and keeps everything of it in shared memory (20 KB), so that no intermediate (per-link
// mask, gradient window, pair list) ever goes to L2 / HBM and the whole stage is a single launch with every tile of the
// pass resident at once:
//   A  coverage  35x35 window of every link's depth plane that reaches into the tile: a warp loads nine rows at a time
//                (lane = column), the coverage bits of a row come straight from the loaded values by ballot; nothing but
//                36 64-bit row masks per link is kept
//   B  pairs     silhouette pixel pairs by XOR of neighbouring row masks (a warp per link, lane = row) -> ONE pair list
//                for the tile
//   C  weights   all threads stride over the tile's list: triangle of the covered pixel (depth plane, L1 / L2 hit), blend
//                weight (ehb_aa_pair)
//   D  masks     link by link through ONE mask buffer: coverage as floats + the four contribution kinds in the reference's
//                order, added to the running sum S in link order
//   E  compose   S = min(sum, 1) -> staging tile -> ONE TMA tensor store (UTMASTG); (S - ref)^2 -> loss; g = dL/dsum of the
//                out region stays in shared memory
//   F  backward  every owned pair with a non-zero weight: analytic gradient of its two edge vertices
//                (ehb_aa_pair_grad), contracted with [x y z 1] on the fly (fp64), reduced per link by shuffles and
//                shared-memory accumulators, 12 fp64 atomics per (tile, link) into d loss / d mvp
// Listed tiles that no triangle reaches (a link's bounding box overlaps them, nothing more) are a zero tile: one TMA store.
// A round handles up to EHB_RL = 8 links of a tile; its pairs live in shared memory (EHB_CAPS) or, for the rare tile with
// more, in a slab of the context's pair pool in global memory.  A tile with more links runs A-D per round for the forward
// and A-C again per round for the backward (g is only known once every link has been composed).
//
// The kernel is compiled for two CTA sizes (this file includes ehb_tiles_impl.cuh twice, namespaces t128 / t256): 128 threads
// for the passes with thousands of listed tiles (more CTAs resident, no second wave), 256 threads for small passes, whose
// duration is the latency of their heaviest tile (640x480, ten views: 78 -> 74 us per solver iteration).
#pragma once
#include "ehb_kernels.cuh"

#define EHB_MROWS (EHB_T + 1)            // out region of a tile: interior + one row / column on the high side
#define EHB_MW 36                        // row pitch (floats) of a link mask / the S / g window
#define EHB_MSZ (EHB_MROWS * EHB_MW)
#ifndef EHB_RL
#define EHB_RL 8                         // links of a tile per round
#endif
#ifndef EHB_CAPS
#define EHB_CAPS 512                     // pairs of a round that fit shared memory
#endif
#define EHB_CAPG 2432                    // pairs of a slab in global memory (>= the 2380 pairs one window can have)
#define EHB_SLAB_BYTES (EHB_CAPG * 12)   // alpha f32 | tri u32 | pk u16 | slot u8 (+ 1 pad)
static_assert(EHB_MSZ * 4 >= EHB_T * EHB_T * 4, "the staging tile of the TMA store reuses the mask buffer");

// EHB_STATS builds: thread 0 of every CTA adds the cycles it spent per phase to the developer counters (ehb_ctx_debug_counters):
// [0] tiles, [1] links, [2] tiles without links, [3] pairs, [4] tiles with several rounds, [5] A windows, [6] B pairs,
// [7] C weights, [8] D masks, [9] E compose, [10] F backward, [11] whole tile, [12] rounds in a global slab
#ifdef EHB_STATS
#define EHB_STAT_T(var) const long long var = clock64()
#define EHB_STAT_ADD(i, v) do { if (threadIdx.x == 0) { atomicAdd(&p.ctr->dbg[i], (unsigned long long)(v)); \
    if ((i) >= 5 && (i) <= 11) sm.stat[(i) - 5] += (long long)(v); if ((i) == 3) sm.stat[7] += (long long)(v); } } while (0)
#else
#define EHB_STAT_T(var)
#define EHB_STAT_ADD(i, v)
#endif

#define EHB_TTHREADS 128
#define EHB_TMIN_BLOCKS 8                // 64 registers: the ~1,100 tiles with links of a 10-view pass are resident at once (+2 % in flight)
#ifndef EHB_TMB128
#define EHB_TMB128 1
#endif
#define EHB_TMB EHB_TMB128                // mask buffers: links whose masks are built together (ehb_tile_masks)
#define EHB_TNS t128
#include "ehb_tiles_impl.cuh"
#undef EHB_TTHREADS
#undef EHB_TMIN_BLOCKS
#undef EHB_TMB
#undef EHB_TNS

#define EHB_TTHREADS 256
#define EHB_TMIN_BLOCKS 3
#define EHB_TMB 6
#define EHB_TNS t256
#include "ehb_tiles_impl.cuh"
#undef EHB_TTHREADS
#undef EHB_TMIN_BLOCKS
#undef EHB_TMB
#undef EHB_TNS
