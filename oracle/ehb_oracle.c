/*
 * ehb_oracle.c -- CPU restatement of EasyHeC's render_mask hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in easyhec_b200/ (the product) may import, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, as the checker and as the timed CPU baseline.
 *
 * PARITY STATUS: the arithmetic of the path lives in nvdiffrast (requirements.txt:29,
 * git+https://github.com/NVlabs/nvdiffrast.git, unpinned, NOT vendored in /root/reference and not
 * installable offline).  The host-side geometry (projection, pose chain, se3) is pinned against
 * the reference's own Python (tests/golden/host_math.npz, tools/make_golden_host_math.py).
 * The rasterize / antialias semantics below restate nvdiffrast's published CUDA algorithm from
 * its call sites in easyhec/structures/nvdiffrast_renderer.py:39-47,64-72 and SURVEY.md
 * Appendix A; there are no reference golden vectors for them => "parity unpinned" for those.
 *
 * What each function follows:
 *   eho_transform            easyhec/utils/nvdiffrast_utils.py:14-18  (transform_pos)
 *   eho_rasterize            dr.rasterize  call  nvdiffrast_renderer.py:39,64  (cudaraster
 *                            triangle setup: 4 sub-pixel bits, round-to-nearest snapping, integer
 *                            edge functions with a tie rule; shader: z/w from fp32 barycentrics)
 *   eho_binary_mask          nvdiffrast_renderer.py:46-47,71-72  (rast[...,2] > 0, row flip)
 *   eho_build_adjacency      dr.antialias topology hash (edge -> opposite vertices)
 *   eho_antialias_fwd/bwd    dr.antialias  call  nvdiffrast_renderer.py:41-44,66-69
 *   eho_render_views         easyhec/modeling/models/rb_solve/rb_solver.py:60-72 (per-link AA mask,
 *                            sum, clamp(max=1), squared error, mean over views) + its backward
 *   eho_union_binary / eho_variance_scores
 *                            easyhec/utils/render_api.py:70-96 + space_explorer.py:152-165
 *
 * PROVENANCE of the antialias arithmetic: the pair analysis (pair_alpha) and its gradient (pair_grad) keep the
 * operation ORDER and the constants of nvdiffrast's published antialias analysis / gradient kernels as summarised
 * in SURVEY.md Appendix A.4 (nvdiffrast itself is NOT under /root/reference and was not available here; its
 * licence is NVIDIA's non-commercial source licence).  The order is parity-mandated -- a different order changes
 * the last bits of the blend weights -- and is the only thing taken over; the rasterizer, the data layout and
 * the work decomposition around it are this repository's own.
 *
 * All fp32 arithmetic is written one rounding per operation (compile with -ffp-contract=off);
 * the CUDA kernels use the same operation order with __fmul_rn/__fadd_rn so that coverage,
 * depth winners and antialias weights are reproducible bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define EHO_API __attribute__((visibility("default")))

/* Tie rule for a sample lying exactly on a snapped edge.
 * With the triangle wound so that its signed area is positive in GL coordinates (y up) and
 * d = (dx,dy) the edge direction, a sample with edge function == 0 is covered iff
 *   rule 0 ("cudaraster recollection"): dy > 0 || (dy == 0 && dx < 0)    right / top edges in GL
 *   rule 1 (mirror):                    dy < 0 || (dy == 0 && dx > 0)    left / bottom edges in GL
 * Interior shared edges are owned by exactly one side under either rule.  Unverifiable without
 * nvdiffrast; isolated here so a golden-vector check can flip it. */
static inline int edge_inclusive(int64_t dx, int64_t dy, int rule)
{
    int r0 = (dy > 0) || (dy == 0 && dx < 0);
    if (dx == 0 && dy == 0) return 0;
    return rule == 0 ? r0 : !r0;
}

static inline int32_t rni_sat(float x)
{
    /* cvt.rni.sat.s32.f32: round to nearest even, saturate, NaN -> 0 */
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)lrintf(x);
}

static inline uint32_t order_key(float f)
{
    uint32_t b;
    memcpy(&b, &f, 4);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

EHO_API int eho_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Thread count of the view-parallel loops below.  Launchers such as torchrun export OMP_NUM_THREADS=1; the timed CPU
 * baseline sets the count explicitly instead of inheriting that. */
EHO_API void eho_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* clip[v] = [x y z 1] * mvp^T, accumulate k = 0..3 the way an fp32 GEMM inner loop does:
 * c = x*m0; c = fma(y,m1,c); c = fma(z,m2,c); c = c + m3. */
EHO_API void eho_transform(const float* verts, int V, const float* mvp, float* clip)
{
    for (int v = 0; v < V; v++) {
        float x = verts[3 * v], y = verts[3 * v + 1], z = verts[3 * v + 2];
        for (int r = 0; r < 4; r++) {
            const float* m = mvp + 4 * r;
            float c = x * m[0];
            c = fmaf(y, m[1], c);
            c = fmaf(z, m[2], c);
            c = c + m[3];
            clip[4 * v + r] = c;
        }
    }
}

/* z/w at pixel centre from unsnapped clip positions (rasterize shader). */
static inline float shade_zw(const float* p0, const float* p1, const float* p2, float fx, float fy)
{
    float p0x = p0[0] - fx * p0[3], p0y = p0[1] - fy * p0[3];
    float p1x = p1[0] - fx * p1[3], p1y = p1[1] - fy * p1[3];
    float p2x = p2[0] - fx * p2[3], p2y = p2[1] - fy * p2[3];
    float a0 = p1x * p2y - p1y * p2x;
    float a1 = p2x * p0y - p2y * p0x;
    float a2 = p0x * p1y - p0y * p1x;
    float z = (p0[2] * a0 + p1[2] * a1) + p2[2] * a2;
    float w = (p0[3] * a0 + p1[3] * a1) + p2[3] * a2;
    float zw = z / w;
    return fminf(fmaxf(zw, -1.f), 1.f);
}

/* Coverage + depth of one triangle given by its SNAPPED vertices (sub-pixel units) into key[]; z/w is shaded from the
 * three ORIGINAL clip-space vertices c0, c1, c2 (the rasterizer's shader recomputes it per pixel from the unclipped
 * positions), id = triangle id. */
static void raster_snapped(int64_t x0, int64_t y0, int64_t x1, int64_t y1, int64_t x2, int64_t y2, const float* c0,
                           const float* c1, const float* c2, uint32_t id, int H, int W, int rule, uint64_t* key)
{
    const float xs = 2.f / (float)W, xo = 1.f / (float)W - 1.f;
    const float ys = 2.f / (float)H, yo = 1.f / (float)H - 1.f;
    int64_t area = (x1 - x0) * (y2 - y0) - (y1 - y0) * (x2 - x0);
    if (area == 0) return;
    if (area < 0) { int64_t tx = x1, ty = y1; x1 = x2; y1 = y2; x2 = tx; y2 = ty; }
    int64_t lox = x0 < x1 ? (x0 < x2 ? x0 : x2) : (x1 < x2 ? x1 : x2);
    int64_t hix = x0 > x1 ? (x0 > x2 ? x0 : x2) : (x1 > x2 ? x1 : x2);
    int64_t loy = y0 < y1 ? (y0 < y2 ? y0 : y2) : (y1 < y2 ? y1 : y2);
    int64_t hiy = y0 > y1 ? (y0 > y2 ? y0 : y2) : (y1 > y2 ? y1 : y2);
    /* samples at 16*p + 8 - 8*W:  p >= ceil((lo + 8W - 8)/16), p <= floor((hi + 8W - 8)/16) */
    int64_t bx = 8 * (int64_t)W - 8, by = 8 * (int64_t)H - 8;
    int64_t pxlo = (lox + bx + 15) >> 4, pxhi = (hix + bx) >> 4;
    int64_t pylo = (loy + by + 15) >> 4, pyhi = (hiy + by) >> 4;
    if (pxlo < 0) pxlo = 0;
    if (pylo < 0) pylo = 0;
    if (pxhi > W - 1) pxhi = W - 1;
    if (pyhi > H - 1) pyhi = H - 1;
    if (pxlo > pxhi || pylo > pyhi) return;
    int64_t ex[3] = {x1 - x0, x2 - x1, x0 - x2}, ey[3] = {y1 - y0, y2 - y1, y0 - y2};
    int64_t ax[3] = {x0, x1, x2}, ay[3] = {y0, y1, y2};
    int64_t thr[3];
    for (int k = 0; k < 3; k++) thr[k] = edge_inclusive(ex[k], ey[k], rule) ? 0 : 1;
    for (int64_t py = pylo; py <= pyhi; py++) {
        int64_t sy = 16 * py - by;
        for (int64_t px = pxlo; px <= pxhi; px++) {
            int64_t sx = 16 * px - bx;
            int in = 1;
            for (int k = 0; k < 3; k++) {
                int64_t e = ex[k] * (sy - ay[k]) - ey[k] * (sx - ax[k]);
                if (e < thr[k]) { in = 0; break; }
            }
            if (!in) continue;
            float fx = xs * (float)px + xo, fy = ys * (float)py + yo;
            float zw = shade_zw(c0, c1, c2, fx, fy);
            uint64_t k64 = ((uint64_t)order_key(zw) << 32) | id;
            uint64_t* dst = key + py * W + px;
            if (k64 < *dst) *dst = k64;
        }
    }
}

/* Sutherland-Hodgman clip of a triangle against the six planes of the view frustum in clip space
 * (x >= -w, x <= w, y >= -w, y <= w, z >= -w, z <= w), fp32, one rounding per operation.  An intersection is always
 * computed from the INSIDE vertex towards the outside one (t = d_in / (d_in - d_out), p = in + t * (out - in)), so an
 * edge shared by two triangles is cut at bit-identical points whichever way each triangle runs through it.
 * out[9][4]; returns the number of vertices (0 when nothing is left). */
static float clip_dist(const float* v, int plane)
{
    switch (plane) {
    case 0: return v[3] + v[0];
    case 1: return v[3] - v[0];
    case 2: return v[3] + v[1];
    case 3: return v[3] - v[1];
    case 4: return v[3] + v[2];
    default: return v[3] - v[2];
    }
}
static int clip_triangle(const float* v0, const float* v1, const float* v2, float out[9][4])
{
    float a[9][4], b[9][4];
    int n = 3;
    memcpy(a[0], v0, 16); memcpy(a[1], v1, 16); memcpy(a[2], v2, 16);
    for (int plane = 0; plane < 6 && n >= 3; plane++) {
        int m = 0;
        for (int i = 0; i < n; i++) {
            const float* p = a[i];
            const float* q = a[(i + 1) % n];
            float dp = clip_dist(p, plane), dq = clip_dist(q, plane);
            int ip = dp >= 0.f, iq = dq >= 0.f;
            if (ip && m < 9) { memcpy(b[m], p, 16); m++; }
            if (ip != iq && m < 9) {
                const float* in = ip ? p : q;
                const float* ou = ip ? q : p;
                float din = ip ? dp : dq, dou = ip ? dq : dp;
                float t = din / (din - dou);
                for (int k = 0; k < 4; k++) b[m][k] = in[k] + t * (ou[k] - in[k]);
                m++;
            }
        }
        n = m;
        memcpy(a, b, sizeof a);
    }
    if (n < 3) return 0;
    memcpy(out, a, sizeof a);
    return n;
}

/* Visibility pass.  key[H*W] (GL rows, row 0 = bottom): (order_key(z/w) << 32) | tri, or ~0 when
 * empty.  A triangle that lies inside the depth range and the fixed-point guard band is snapped and drawn directly;
 * any other one that survives the trivial frustum rejection is clipped against the view frustum first and drawn as a
 * fan of sub-triangles with its own id and its own (unclipped) vertices for the depth -- like cudaraster's triangle
 * setup.  Returns the number of triangles that went through the clipper. */
EHO_API int eho_rasterize(const float* clip, int V, const int* tri, int F, int H, int W, int rule,
                          uint64_t* key)
{
    int nclip = 0;
    const float vsx = (float)(W * 8), vsy = (float)(H * 8);
    for (long i = 0; i < (long)H * W; i++) key[i] = ~0ull;
    for (int t = 0; t < F; t++) {
        int i0 = tri[3 * t], i1 = tri[3 * t + 1], i2 = tri[3 * t + 2];
        if (i0 < 0 || i0 >= V || i1 < 0 || i1 >= V || i2 < 0 || i2 >= V) continue;
        const float *v0 = clip + 4 * i0, *v1 = clip + 4 * i1, *v2 = clip + 4 * i2;
        /* all three outside one frustum plane => culled */
        if ((v0[3] < v0[0] && v1[3] < v1[0] && v2[3] < v2[0]) || (v0[3] < -v0[0] && v1[3] < -v1[0] && v2[3] < -v2[0]) ||
            (v0[3] < v0[1] && v1[3] < v1[1] && v2[3] < v2[1]) || (v0[3] < -v0[1] && v1[3] < -v1[1] && v2[3] < -v2[1]) ||
            (v0[3] < v0[2] && v1[3] < v1[2] && v2[3] < v2[2]) || (v0[3] < -v0[2] && v1[3] < -v1[2] && v2[3] < -v2[2]))
            continue;
        int direct = v0[3] >= fabsf(v0[2]) && v1[3] >= fabsf(v1[2]) && v2[3] >= fabsf(v2[2]);
        int64_t x0 = 0, y0 = 0, x1 = 0, y1 = 0, x2 = 0, y2 = 0;
        if (direct) {
            float r0 = 1.0f / v0[3], r1 = 1.0f / v1[3], r2 = 1.0f / v2[3];
            x0 = rni_sat(v0[0] * r0 * vsx); y0 = rni_sat(v0[1] * r0 * vsy);
            x1 = rni_sat(v1[0] * r1 * vsx); y1 = rni_sat(v1[1] * r1 * vsy);
            x2 = rni_sat(v2[0] * r2 * vsx); y2 = rni_sat(v2[1] * r2 * vsy);
            /* guard band: beyond +-2^28 sub-pixel units the 64-bit edge products could overflow: clipped instead */
            const int64_t G = (int64_t)1 << 28;
            if (x0 > G || x0 < -G || y0 > G || y0 < -G || x1 > G || x1 < -G || y1 > G || y1 < -G ||
                x2 > G || x2 < -G || y2 > G || y2 < -G) direct = 0;
        }
        if (direct) {
            raster_snapped(x0, y0, x1, y1, x2, y2, v0, v1, v2, (uint32_t)t, H, W, rule, key);
            continue;
        }
        nclip++;
        float poly[9][4];
        int n = clip_triangle(v0, v1, v2, poly);
        int64_t sx[9], sy[9];
        int ok = 1;
        for (int i = 0; i < n; i++) {
            if (!(poly[i][3] > 0.f)) { ok = 0; break; }
            float r = 1.0f / poly[i][3];
            sx[i] = rni_sat(poly[i][0] * r * vsx);
            sy[i] = rni_sat(poly[i][1] * r * vsy);
        }
        if (!ok) continue;
        for (int i = 1; i + 1 < n; i++)
            raster_snapped(sx[0], sy[0], sx[i], sy[i], sx[i + 1], sy[i + 1], v0, v1, v2, (uint32_t)t, H, W, rule, key);
    }
    return nclip;
}

/* rast[...,2] > 0 then torch.flip(dims=[0]): out[r][c] = (z/w of nearest > 0), r = H-1-py. */
EHO_API void eho_binary_mask(const uint64_t* key, int H, int W, uint8_t* out)
{
    for (int py = 0; py < H; py++)
        for (int px = 0; px < W; px++) {
            uint64_t k = key[(long)py * W + px];
            out[(long)(H - 1 - py) * W + px] = (k != ~0ull) && ((uint32_t)(k >> 32) > 0x80000000u);
        }
}

/* ---------------------------------------------------------------------------------------------
 * Topology: for every triangle t and corner k, opp[3t+k] = the vertex opposite to edge k
 * (edge k joins corners k+1 and k+2) in the *other* triangle sharing that edge, or -1.
 * Mirrors the antialias edge hash: an edge remembers the first two opposite vertices inserted
 * (triangle order here; insertion order is a race in the original). */
typedef struct { int a, b, n0, n1; } edge_rec;
static int edge_cmp(const void* p, const void* q)
{
    const int* x = (const int*)p; const int* y = (const int*)q;
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    if (x[1] != y[1]) return x[1] < y[1] ? -1 : 1;
    return x[2] < y[2] ? -1 : (x[2] > y[2]);
}
EHO_API void eho_build_adjacency(const int* tri, int F, int V, int* opp)
{
    /* records: (min, max, order, opposite) sorted by (min,max,order) */
    int* rec = (int*)malloc(sizeof(int) * 4 * 3 * (size_t)F);
    int n = 0;
    for (int t = 0; t < F; t++) {
        int v[3] = {tri[3 * t], tri[3 * t + 1], tri[3 * t + 2]};
        int bad = 0;
        for (int k = 0; k < 3; k++) if (v[k] < 0 || v[k] >= V) bad = 1;
        if (bad || v[0] == v[1] || v[1] == v[2] || v[2] == v[0]) continue;
        for (int k = 0; k < 3; k++) {
            int a = v[(k + 1) % 3], b = v[(k + 2) % 3];
            rec[4 * n] = a < b ? a : b; rec[4 * n + 1] = a < b ? b : a; rec[4 * n + 2] = 3 * t + k; rec[4 * n + 3] = v[k];
            n++;
        }
    }
    qsort(rec, n, 4 * sizeof(int), edge_cmp);
    for (int i = 0; i < 3 * F; i++) opp[i] = -1;
    for (int i = 0; i < n;) {
        int j = i;
        while (j < n && rec[4 * j] == rec[4 * i] && rec[4 * j + 1] == rec[4 * i + 1]) j++;
        int n0 = rec[4 * i + 3], n1 = (j - i >= 2) ? rec[4 * (i + 1) + 3] : -1;
        for (int k = i; k < j; k++) {
            int vr = rec[4 * k + 3];
            opp[rec[4 * k + 2]] = (n0 == vr) ? n1 : n0;
        }
        i = j;
    }
    free(rec);
}

/* ---------------------------------------------------------------------------------------------
 * Antialias.  Colour is exactly 1 on covered pixels and 0 elsewhere (interpolate of a ones
 * attribute), so only covered<->empty pixel pairs change the image or carry gradient. */
static inline int same_sign(float a, float b)
{
    int32_t ia, ib; memcpy(&ia, &a, 4); memcpy(&ib, &b, 4);
    return (ia ^ ib) >= 0;
}
static inline int rational_gt(float n0, float n1, float d0, float d1) { return (n0 * d1 > n1 * d0) == same_sign(d0, d1); }
static inline int max_idx3(float n0, float n1, float n2, float d0, float d1, float d2)
{
    int g10 = rational_gt(n1, n0, d1, d0), g20 = rational_gt(n2, n0, d2, d0), g21 = rational_gt(n2, n1, d2, d1);
    if (g20 && g21) return 2;
    if (g10) return 1;
    return 0;
}
#define EHO_F32_MAX 3.402823466e+38f

/* One pixel pair.  (px,py) = p0 in GL rows, d = 0: p1 = right neighbour, d = 1: p1 = row above (py+1).
 * Returns alpha (0 when no edge found) and the edge index / side in *di,*side (side 1: triangle is p1's). */
static float aa_pair(const float* clip, const int* tri, const int* opp, int t, int side, int px, int py, int d,
                     int H, int W, int* di_out)
{
    const float xh = 0.5f * (float)W, yh = 0.5f * (float)H;
    if (side) { px += 1 - d; py += d; }
    int vi0 = tri[3 * t], vi1 = tri[3 * t + 1], vi2 = tri[3 * t + 2];
    int o0 = opp[3 * t], o1 = opp[3 * t + 1], o2 = opp[3 * t + 2];
    const float *p0 = clip + 4 * vi0, *p1 = clip + 4 * vi1, *p2 = clip + 4 * vi2;
    const float *q0 = o0 < 0 ? p0 : clip + 4 * o0, *q1 = o1 < 0 ? p1 : clip + 4 * o1, *q2 = o2 < 0 ? p2 : clip + 4 * o2;
    float w0 = 1.f / p0[3], w1 = 1.f / p1[3], w2 = 1.f / p2[3];
    float ow0 = 1.f / q0[3], ow1 = 1.f / q1[3], ow2 = 1.f / q2[3];
    float fx = (float)px + .5f - xh, fy = (float)py + .5f - yh;
    float x0 = p0[0] * w0 * xh - fx, y0 = p0[1] * w0 * yh - fy;
    float x1 = p1[0] * w1 * xh - fx, y1 = p1[1] * w1 * yh - fy;
    float x2 = p2[0] * w2 * xh - fx, y2 = p2[1] * w2 * yh - fy;
    float ox0 = q0[0] * ow0 * xh - fx, oy0 = q0[1] * ow0 * yh - fy;
    float ox1 = q1[0] * ow1 * xh - fx, oy1 = q1[1] * ow1 * yh - fy;
    float ox2 = q2[0] * ow2 * xh - fx, oy2 = q2[1] * ow2 * yh - fy;
    float bb = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0);
    float a0 = (x1 - ox0) * (y2 - oy0) - (x2 - ox0) * (y1 - oy0);
    float a1 = (x2 - ox1) * (y0 - oy1) - (x0 - ox1) * (y2 - oy1);
    float a2 = (x0 - ox2) * (y1 - oy2) - (x1 - ox2) * (y0 - oy2);
    *di_out = 0;
    if (!(same_sign(a0, bb) || same_sign(a1, bb) || same_sign(a2, bb))) return 0.f;
    if (d) { float s; s = x0; x0 = y0; y0 = s; s = x1; x1 = y1; y1 = s; s = x2; x2 = y2; y2 = s; }
    float dx0 = x2 - x1, dx1 = x0 - x2, dx2 = x1 - x0;
    float dy0 = y2 - y1, dy1 = y0 - y2, dy2 = y1 - y0;
    float dc = -EHO_F32_MAX;
    float ds = side ? -1.f : 1.f;
    float d0 = ds * (x1 * dy0 - y1 * dx0);
    float d1 = ds * (x2 * dy1 - y2 * dx1);
    float d2 = ds * (x0 * dy2 - y0 * dx2);
    if (same_sign(y1, y2)) { d0 = -EHO_F32_MAX; dy0 = 1.f; }
    if (same_sign(y2, y0)) { d1 = -EHO_F32_MAX; dy1 = 1.f; }
    if (same_sign(y0, y1)) { d2 = -EHO_F32_MAX; dy2 = 1.f; }
    int di = max_idx3(d0, d1, d2, dy0, dy1, dy2);
    if (di == 0 && same_sign(a0, bb) && fabsf(dy0) >= fabsf(dx0)) dc = d0 / dy0;
    if (di == 1 && same_sign(a1, bb) && fabsf(dy1) >= fabsf(dx1)) dc = d1 / dy1;
    if (di == 2 && same_sign(a2, bb) && fabsf(dy2) >= fabsf(dx2)) dc = d2 / dy2;
    const float eps = .0625f;
    if (dc > -eps && dc < 1.f + eps) {
        dc = fminf(fmaxf(dc, 0.f), 1.f);
        *di_out = di;
        return ds * (.5f - dc);
    }
    return 0.f;
}

/* Forward.  out_gl[H*W] (GL rows) = colour + blends.  alpha[2*H*W]: alpha[d*H*W + p0] per pair, and
 * info[2*H*W] = di | side<<2 (valid where alpha != 0).  Per-pixel accumulation order is fixed:
 * colour, pair(p,p+x), pair(p,p+y), pair(p-x,p), pair(p-y,p). */
EHO_API void eho_antialias_fwd(const float* clip, const int* tri, const int* opp, const uint64_t* key,
                               int H, int W, float* out_gl, float* alpha, uint8_t* info)
{
    long n = (long)H * W;
    memset(alpha, 0, sizeof(float) * 2 * n);
    memset(info, 0, 2 * n);
    for (int py = 0; py < H; py++)
        for (int px = 0; px < W; px++) {
            long p0 = (long)py * W + px;
            int c0 = key[p0] != ~0ull;
            for (int d = 0; d < 2; d++) {
                if (d == 0 && px >= W - 1) continue;
                if (d == 1 && py >= H - 1) continue;
                long p1 = p0 + (d ? W : 1);
                int c1 = key[p1] != ~0ull;
                if (c0 == c1) continue;   /* both covered: colour difference is 0; both empty: no item */
                int side = c0 ? 0 : 1;
                int t = (int)(uint32_t)(c0 ? key[p0] : key[p1]);
                int di;
                float a = aa_pair(clip, tri, opp, t, side, px, py, d, H, W, &di);
                alpha[d * n + p0] = a;
                info[d * n + p0] = (uint8_t)(di | (side << 2));
            }
        }
    for (int py = 0; py < H; py++)
        for (int px = 0; px < W; px++) {
            long p = (long)py * W + px;
            float c = key[p] != ~0ull ? 1.f : 0.f;
            float o = c;
            /* pair (p, p+x): p is p0 -> receives when alpha > 0 ; colour diff = c1 - c0 */
            if (px < W - 1) { float a = alpha[p]; if (a > 0.f) o += a * ((key[p + 1] != ~0ull ? 1.f : 0.f) - c); }
            if (py < H - 1) { float a = alpha[n + p]; if (a > 0.f) o += a * ((key[p + W] != ~0ull ? 1.f : 0.f) - c); }
            /* pair (p-x, p): p is p1 -> receives when alpha <= 0 (alpha == 0 adds nothing) */
            if (px > 0) { float a = alpha[p - 1]; if (!(a > 0.f) && a != 0.f) o += a * (c - (key[p - 1] != ~0ull ? 1.f : 0.f)); }
            if (py > 0) { float a = alpha[n + p - W]; if (!(a > 0.f) && a != 0.f) o += a * (c - (key[p - W] != ~0ull ? 1.f : 0.f)); }
            out_gl[p] = o;
        }
}

/* Backward of the antialias blend w.r.t. clip-space positions.  dy_gl[H*W] = dL/d(out_gl).
 * gpos[V*4] (double, accumulated: caller zeroes) receives x,y,w components. */
EHO_API void eho_antialias_bwd(const float* clip, const int* tri, const uint64_t* key, int H, int W,
                               const float* alpha, const uint8_t* info, const float* dy_gl, double* gpos)
{
    long n = (long)H * W;
    for (int d = 0; d < 2; d++)
        for (int py0 = 0; py0 < H; py0++)
            for (int px0 = 0; px0 < W; px0++) {
                long p0 = (long)py0 * W + px0;
                float al = alpha[d * n + p0];
                if (al == 0.f) continue;
                long p1 = p0 + (d ? W : 1);
                int di = info[d * n + p0] & 3, side = (info[d * n + p0] >> 2) & 1;
                int t = (int)(uint32_t)(side ? key[p1] : key[p0]);
                int px = px0, py = py0;
                if (side) { px += 1 - d; py += d; }
                float g = dy_gl[al > 0.f ? p0 : p1];
                float c0 = key[p0] != ~0ull ? 1.f : 0.f, c1 = key[p1] != ~0ull ? 1.f : 0.f;
                float dd = 0.f;
                if (g != 0.f) dd += g * (c1 - c0);
                if (dd == 0.f) continue;
                int i1 = (di < 2) ? (di + 1) : 0, i2 = (i1 < 2) ? (i1 + 1) : 0;
                int vi1 = tri[3 * t + i1], vi2 = tri[3 * t + i2];
                float p1v[4], p2v[4];
                memcpy(p1v, clip + 4 * vi1, 16); memcpy(p2v, clip + 4 * vi2, 16);
                float pxh = 0.5f * (float)W, pyh = 0.5f * (float)H;
                float fx = (float)px + .5f - pxh, fy = (float)py + .5f - pyh;
                if (d) { float s; s = p1v[0]; p1v[0] = p1v[1]; p1v[1] = s; s = p2v[0]; p2v[0] = p2v[1]; p2v[1] = s;
                         s = pxh; pxh = pyh; pyh = s; s = fx; fx = fy; fy = s; }
                float w1 = 1.f / p1v[3], w2 = 1.f / p2v[3];
                float x1 = p1v[0] * w1 * pxh - fx, y1 = p1v[1] * w1 * pyh - fy;
                float x2 = p2v[0] * w2 * pxh - fx, y2 = p2v[1] * w2 * pyh - fy;
                float dx = x2 - x1, dy = y2 - y1;
                float db = x1 * dy - y1 * dx;
                float ep = copysignf(1e-3f, dy);
                float iy = 1.f / (dy + ep);
                float dby = db * iy;
                float iw1 = -w1 * iy * dd, iw2 = w2 * iy * dd;
                float gp1x = iw1 * pxh * y2, gp2x = iw2 * pxh * y1;
                float gp1y = iw1 * pyh * (dby - x2), gp2y = iw2 * pyh * (dby - x1);
                float gp1w = -(p1v[0] * gp1x + p1v[1] * gp1y) * w1;
                float gp2w = -(p2v[0] * gp2x + p2v[1] * gp2y) * w2;
                if (d) { float s; s = gp1x; gp1x = gp1y; gp1y = s; s = gp2x; gp2x = gp2y; gp2y = s; }
                if (fabsf(al) >= 0.5f) { gp1x = gp1y = gp1w = 0.f; gp2x = gp2y = gp2w = 0.f; }
                gpos[4 * vi1 + 0] += gp1x; gpos[4 * vi1 + 1] += gp1y; gpos[4 * vi1 + 3] += gp1w;
                gpos[4 * vi2 + 0] += gp2x; gpos[4 * vi2 + 1] += gp2y; gpos[4 * vi2 + 3] += gp2w;
            }
}

/* g_mvp[r][c] = sum_v gpos[v][r] * [x y z 1][c]   (backward of eho_transform w.r.t. mvp) */
static void gpos_to_gmvp(const double* gpos, const float* verts, int V, double* gmvp)
{
    for (int i = 0; i < 16; i++) gmvp[i] = 0.0;
    for (int v = 0; v < V; v++) {
        double h[4] = {verts[3 * v], verts[3 * v + 1], verts[3 * v + 2], 1.0};
        for (int r = 0; r < 4; r++) {
            double g = gpos[4 * v + r];
            if (g == 0.0) continue;
            for (int c = 0; c < 4; c++) gmvp[4 * r + c] += g * h[c];
        }
    }
}

/* ---------------------------------------------------------------------------------------------
 * The single-mesh operator: render_mask(verts, faces, K, pose) with mvp = proj @ flip @ pose.
 * aa = 1: out_f32[H*W] float mask (image rows, row 0 = top); aa = 0: out_u8[H*W] bool mask.
 * Optional saved state for the backward: key / alpha / info (caller-allocated, may be NULL). */
EHO_API int eho_render_mask(const float* verts, int V, const int* tri, int F, const int* opp, const float* mvp,
                            int H, int W, int aa, int rule, float* out_f32, uint8_t* out_u8,
                            uint64_t* key_save, float* alpha_save, uint8_t* info_save)
{
    long n = (long)H * W;
    float* clip = (float*)malloc(sizeof(float) * 4 * (size_t)(V > 0 ? V : 1));
    uint64_t* key = key_save ? key_save : (uint64_t*)malloc(8 * n);
    eho_transform(verts, V, mvp, clip);
    int nclip = eho_rasterize(clip, V, tri, F, H, W, rule, key);
    if (!aa) {
        eho_binary_mask(key, H, W, out_u8);
    } else {
        float* gl = (float*)malloc(sizeof(float) * n);
        float* alpha = alpha_save ? alpha_save : (float*)malloc(sizeof(float) * 2 * n);
        uint8_t* info = info_save ? info_save : (uint8_t*)malloc(2 * n);
        eho_antialias_fwd(clip, tri, opp, key, H, W, gl, alpha, info);
        for (int r = 0; r < H; r++) memcpy(out_f32 + (long)r * W, gl + (long)(H - 1 - r) * W, sizeof(float) * W);
        free(gl);
        if (!alpha_save) free(alpha);
        if (!info_save) free(info);
    }
    if (!key_save) free(key);
    free(clip);
    return nclip;
}

/* Backward of eho_render_mask (aa = 1): dy[H*W] in image rows -> gpos[V*4] (double) and gmvp[16]. */
EHO_API void eho_render_mask_bwd(const float* verts, int V, const int* tri, int F, const float* mvp, int H, int W,
                                 const uint64_t* key, const float* alpha, const uint8_t* info, const float* dy,
                                 double* gpos, double* gmvp)
{
    (void)F;
    long n = (long)H * W;
    float* clip = (float*)malloc(sizeof(float) * 4 * (size_t)(V > 0 ? V : 1));
    float* dgl = (float*)malloc(sizeof(float) * n);
    eho_transform(verts, V, mvp, clip);
    for (int r = 0; r < H; r++) memcpy(dgl + (long)(H - 1 - r) * W, dy + (long)r * W, sizeof(float) * W);
    memset(gpos, 0, sizeof(double) * 4 * (size_t)V);
    eho_antialias_bwd(clip, tri, key, H, W, alpha, info, dgl, gpos);
    gpos_to_gmvp(gpos, verts, V, gmvp);
    free(clip); free(dgl);
}

/* ---------------------------------------------------------------------------------------------
 * RBSolver mask loop, forward + backward, B views x L links (rb_solver.py:60-72):
 *   S_b = min(sum_l aa_mask_{b,l}, 1);  loss = (1/B) sum_b sum_px (S_b - ref_b)^2
 * Geometry is given as L meshes (concatenated arrays with offsets).  Outputs:
 *   masks[B*H*W] (image rows), loss_b[B] (per-view sum of squares, double), gmvp[B*L*16] (double,
 *   d loss / d mvp[b,l] for upstream grad 1).  do_bwd = 0 skips the gradient.  OpenMP over views. */
EHO_API int eho_render_views(int B, int L, const float* verts, const int* voff, const int* tri, const int* opp,
                             const int* foff, const float* mvp, const float* ref, int H, int W, int rule, int do_bwd,
                             float* masks, double* loss_b, double* gmvp)
{
    long n = (long)H * W;
    int Vmax = 0, total_clip = 0;
    for (int l = 0; l < L; l++) if (voff[l + 1] - voff[l] > Vmax) Vmax = voff[l + 1] - voff[l];
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total_clip)
    for (int b = 0; b < B; b++) {
        float* clip = (float*)malloc(sizeof(float) * 4 * (size_t)(voff[L] > 0 ? voff[L] : 1));
        uint64_t* key = (uint64_t*)malloc(8 * n * (size_t)L);
        float* alpha = (float*)malloc(sizeof(float) * 2 * n * (size_t)L);
        uint8_t* info = (uint8_t*)malloc(2 * n * (size_t)L);
        float* gl = (float*)malloc(sizeof(float) * n);
        float* sum = (float*)malloc(sizeof(float) * n);
        float* g = (float*)malloc(sizeof(float) * n);
        double* gpos = (double*)malloc(sizeof(double) * 4 * (size_t)(Vmax > 0 ? Vmax : 1));
        for (int l = 0; l < L; l++) {
            int V = voff[l + 1] - voff[l], F = foff[l + 1] - foff[l];
            const float* m = mvp + ((long)b * L + l) * 16;
            float* cl = clip + 4 * (long)voff[l];
            eho_transform(verts + 3 * (long)voff[l], V, m, cl);
            total_clip += eho_rasterize(cl, V, tri + 3 * (long)foff[l], F, H, W, rule, key + n * l);
            eho_antialias_fwd(cl, tri + 3 * (long)foff[l], opp + 3 * (long)foff[l], key + n * l, H, W, gl,
                              alpha + 2 * n * l, info + 2 * n * l);
            if (l == 0) memcpy(sum, gl, sizeof(float) * n);
            else for (long i = 0; i < n; i++) sum[i] = sum[i] + gl[i];
        }
        double loss = 0.0;
        const float invB = 1.0f / (float)B;
        for (int py = 0; py < H; py++)
            for (int px = 0; px < W; px++) {
                long pg = (long)py * W + px, pi = (long)(H - 1 - py) * W + px;
                float s = sum[pg];
                float S = s > 1.f ? 1.f : s;
                masks[(long)b * n + pi] = S;
                float diff = S - ref[(long)b * n + pi];
                loss += (double)(diff * diff);
                g[pg] = (s <= 1.f) ? (2.f * diff) * invB : 0.f;   /* clamp(max=1) passes grad where sum <= 1 */
            }
        loss_b[b] = loss;
        if (do_bwd)
            for (int l = 0; l < L; l++) {
                int V = voff[l + 1] - voff[l];
                memset(gpos, 0, sizeof(double) * 4 * (size_t)V);
                eho_antialias_bwd(clip + 4 * (long)voff[l], tri + 3 * (long)foff[l], key + n * l, H, W,
                                  alpha + 2 * n * l, info + 2 * n * l, g, gpos);
                gpos_to_gmvp(gpos, verts + 3 * (long)voff[l], V, gmvp + ((long)b * L + l) * 16);
            }
        free(clip); free(key); free(alpha); free(info); free(gl); free(sum); free(g); free(gpos);
    }
    return total_clip;
}

/* ---------------------------------------------------------------------------------------------
 * Space exploration inner loop (render_api.py:70-96 -> batch_render_mask(anti_aliasing=False),
 * space_explorer.py:163-164).  N renders of the packed robot: L links drawn into ONE depth buffer
 * with mvp[n,l]; out[n] = bool mask.  Triangle ids are offset by the link's first face so ties
 * break like a packed mesh. */
EHO_API int eho_union_binary(int N, int L, const float* verts, const int* voff, const int* tri, const int* foff,
                             const float* mvp, int H, int W, int rule, uint8_t* out)
{
    long n = (long)H * W;
    int total_clip = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total_clip)
    for (int i = 0; i < N; i++) {
        int Vt = voff[L], Ft = foff[L];
        float* clip = (float*)malloc(sizeof(float) * 4 * (size_t)(Vt > 0 ? Vt : 1));
        int* ptri = (int*)malloc(sizeof(int) * 3 * (size_t)(Ft > 0 ? Ft : 1));
        uint64_t* key = (uint64_t*)malloc(8 * n);
        for (int l = 0; l < L; l++) {
            eho_transform(verts + 3 * (long)voff[l], voff[l + 1] - voff[l], mvp + ((long)i * L + l) * 16, clip + 4 * (long)voff[l]);
            for (int f = foff[l]; f < foff[l + 1]; f++)
                for (int k = 0; k < 3; k++) ptri[3 * f + k] = tri[3 * f + k] + voff[l];
        }
        total_clip += eho_rasterize(clip, Vt, ptri, Ft, H, W, rule, key);
        eho_binary_mask(key, H, W, out + (long)i * n);
        free(clip); free(ptri); free(key);
    }
    return total_clip;
}

/* masks[Q*C*H*W] bool -> score[q] = sum_px var_c(mask) (unbiased, torch.var default).  For 0/1 samples the unbiased
 * variance of a pixel with k ones out of C is k (C - k) / (C (C - 1)), so the sum over pixels is an integer divided by
 * C (C - 1): accumulated exactly (the reference sums fp32 variances; its value carries ~1e-7 relative rounding). */
EHO_API void eho_variance_scores(const uint8_t* masks, int Q, int C, long n, double* score)
{
    for (int q = 0; q < Q; q++) {
        unsigned long long acc = 0;
        for (long p = 0; p < n; p++) {
            int k = 0;
            for (int c = 0; c < C; c++) k += masks[((long)q * C + c) * n + p] != 0;
            acc += (unsigned long long)(k * (C - k));
        }
        score[q] = C > 1 ? (double)acc / ((double)C * (double)(C - 1)) : 0.0;
    }
}
