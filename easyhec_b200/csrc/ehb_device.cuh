// ehb_device.cuh -- per-triangle and per-pixel-pair device math of the silhouette rasterizer.
//
// Semantics (what must come out) follow the render_mask operator of EasyHeC
// (easyhec/structures/nvdiffrast_renderer.py:25-48: rasterize -> interpolate(ones) -> antialias ->
// channel 0 -> row flip) as specified in SURVEY.md Appendix A.  The arithmetic is written one IEEE
// fp32 rounding per operation in a fixed order (this translation unit is compiled with -fmad=false;
// intended fused operations use __fmaf_rn explicitly) so that coverage, depth winners and the
// antialias weights are reproducible bit for bit against the CPU checker in oracle/.
//
// Provenance of the antialias arithmetic: ehb_aa_pair / ehb_aa_pair_grad keep the operation ORDER and the constants of
// nvdiffrast's published antialias analysis / gradient kernels as summarised in SURVEY.md Appendix A.4 (nvdiffrast is
// not vendored under /root/reference and was not available here; NVIDIA non-commercial source licence).  The order is
// parity-mandated -- another order changes the last bits of the blend weights -- and is the only thing taken over: the
// rasterizer, the data layout and the work decomposition are this repository's own (no bin / coarse / fine pipeline).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define EHB_MAX_LINKS 32
#define EHB_FACE_MASK ((1u << 26) - 1u)   // triangle ids of a link fit in 26 bits
#define EHB_EMPTY 0xFFFFFFFFFFFFFFFFull
#define EHB_F32_MAX 3.402823466e+38f

struct EhbLink {
    const float4* verts;  // [V]  xyz, w unused (1)
    const int4* faces;    // [F]  i0 i1 i2, w unused
    const int4* opp;      // [F]  vertex opposite to edge k in the neighbouring triangle, or -1
    const float4* boxes;  // [2*nboxes] object-space AABBs (min, max) of nboxes contiguous vertex chunks
    const float4* fboxes; // [2*ceil(F/32)] object-space AABBs of the batches of 32 consecutive faces (k_raster's unit of work)
    int V, F, nboxes, pad;
};

struct EhbRobot {
    EhbLink link[EHB_MAX_LINKS];
    int foff[EHB_MAX_LINKS + 1];  // prefix sum of F over links
    int voff[EHB_MAX_LINKS + 1];  // prefix sum of V over links
    int boff[EHB_MAX_LINKS + 1];  // prefix sum of ceil(F / 32) over links: a batch of k_raster never straddles two links
    int L;
};

__device__ __forceinline__ void ehb_xform(const float4 v, const float* __restrict__ m, float* c)
{
    // clip[r] = x*m[r][0] (+) y*m[r][1] (+) z*m[r][2] + m[r][3]; first three fused like a GEMM inner loop
#pragma unroll
    for (int r = 0; r < 4; r++) {
        float a = v.x * m[4 * r];
        a = __fmaf_rn(v.y, m[4 * r + 1], a);
        a = __fmaf_rn(v.z, m[4 * r + 2], a);
        c[r] = a + m[4 * r + 3];
    }
}

__device__ __forceinline__ int ehb_rni_sat(float x)
{
    int r;
    asm("cvt.rni.sat.s32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ uint32_t ehb_order_key(float f)
{
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// Tie rule for a sample exactly on a snapped edge (dx,dy = direction of the counter-clockwise edge).
__device__ __forceinline__ int ehb_edge_inclusive(int dx, int dy, int rule)
{
    if (dx == 0 && dy == 0) return 0;
    const int r0 = (dy > 0) || (dy == 0 && dx < 0);
    return rule == 0 ? r0 : !r0;
}

// z/w at a pixel centre from the unsnapped clip positions, clamped to [-1,1].
__device__ __forceinline__ float ehb_shade_zw(const float* p0, const float* p1, const float* p2, float fx, float fy)
{
    const float p0x = p0[0] - fx * p0[3], p0y = p0[1] - fy * p0[3];
    const float p1x = p1[0] - fx * p1[3], p1y = p1[1] - fy * p1[3];
    const float p2x = p2[0] - fx * p2[3], p2y = p2[1] - fy * p2[3];
    const float a0 = p1x * p2y - p1y * p2x;
    const float a1 = p2x * p0y - p2y * p0x;
    const float a2 = p0x * p1y - p0y * p1x;
    const float z = (p0[2] * a0 + p1[2] * a1) + p2[2] * a2;
    const float w = (p0[3] * a0 + p1[3] * a1) + p2[3] * a2;
    const float zw = z / w;
    return fminf(fmaxf(zw, -1.f), 1.f);
}

// ---------------------------------------------------------------------------------------------------------
// antialias: one horizontally (d=0) or vertically (d=1) adjacent pixel pair, p0 = (px,py) in GL rows.

__device__ __forceinline__ int ehb_same_sign(float a, float b) { return (__float_as_int(a) ^ __float_as_int(b)) >= 0; }
__device__ __forceinline__ int ehb_rational_gt(float n0, float n1, float d0, float d1)
{
    return (n0 * d1 > n1 * d0) == ehb_same_sign(d0, d1);
}
__device__ __forceinline__ int ehb_max_idx3(float n0, float n1, float n2, float d0, float d1, float d2)
{
    const int g10 = ehb_rational_gt(n1, n0, d1, d0), g20 = ehb_rational_gt(n2, n0, d2, d0),
              g21 = ehb_rational_gt(n2, n1, d2, d1);
    if (g20 && g21) return 2;
    if (g10) return 1;
    return 0;
}

// Blend weight of the pair whose covered pixel shows triangle t (side 0: p0 is the covered one, 1: p1).
// Returns alpha (0 = no silhouette edge crosses the segment between the two centres); *di_out = edge index.
// `vc` = clip-space positions of the link's vertices for this item, as written by k_front (same ehb_xform, same matrix:
// bit-identical to transforming the vertex again).
__device__ __forceinline__ void ehb_clip_of(const float4* __restrict__ vc, int v, float* c)
{
    const float4 q = vc[v];
    c[0] = q.x; c[1] = q.y; c[2] = q.z; c[3] = q.w;
}

__device__ __noinline__ float ehb_aa_pair(const EhbLink& lk, const float4* __restrict__ vc, int t, int side, int px,
                                          int py, int d, int H, int W, int* di_out)
{
    const float xh = 0.5f * (float)W, yh = 0.5f * (float)H;
    if (side) { px += 1 - d; py += d; }
    const int4 vi = __ldg(lk.faces + t);
    const int4 oi = __ldg(lk.opp + t);
    float p0[4], p1[4], p2[4], q0[4], q1[4], q2[4];
    ehb_clip_of(vc, vi.x, p0);
    ehb_clip_of(vc, vi.y, p1);
    ehb_clip_of(vc, vi.z, p2);
    ehb_clip_of(vc, oi.x < 0 ? vi.x : oi.x, q0);
    ehb_clip_of(vc, oi.y < 0 ? vi.y : oi.y, q1);
    ehb_clip_of(vc, oi.z < 0 ? vi.z : oi.z, q2);
    const float w0 = 1.f / p0[3], w1 = 1.f / p1[3], w2 = 1.f / p2[3];
    const float ow0 = 1.f / q0[3], ow1 = 1.f / q1[3], ow2 = 1.f / q2[3];
    const float fx = (float)px + .5f - xh, fy = (float)py + .5f - yh;
    float x0 = p0[0] * w0 * xh - fx, y0 = p0[1] * w0 * yh - fy;
    float x1 = p1[0] * w1 * xh - fx, y1 = p1[1] * w1 * yh - fy;
    float x2 = p2[0] * w2 * xh - fx, y2 = p2[1] * w2 * yh - fy;
    const float ox0 = q0[0] * ow0 * xh - fx, oy0 = q0[1] * ow0 * yh - fy;
    const float ox1 = q1[0] * ow1 * xh - fx, oy1 = q1[1] * ow1 * yh - fy;
    const float ox2 = q2[0] * ow2 * xh - fx, oy2 = q2[1] * ow2 * yh - fy;
    const float bb = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0);
    const float a0 = (x1 - ox0) * (y2 - oy0) - (x2 - ox0) * (y1 - oy0);
    const float a1 = (x2 - ox1) * (y0 - oy1) - (x0 - ox1) * (y2 - oy1);
    const float a2 = (x0 - ox2) * (y1 - oy2) - (x1 - ox2) * (y0 - oy2);
    *di_out = 0;
    if (!(ehb_same_sign(a0, bb) || ehb_same_sign(a1, bb) || ehb_same_sign(a2, bb))) return 0.f;
    if (d) { float s; s = x0; x0 = y0; y0 = s; s = x1; x1 = y1; y1 = s; s = x2; x2 = y2; y2 = s; }
    const float dx0 = x2 - x1, dx1 = x0 - x2, dx2 = x1 - x0;
    float dy0 = y2 - y1, dy1 = y0 - y2, dy2 = y1 - y0;
    float dc = -EHB_F32_MAX;
    const float ds = side ? -1.f : 1.f;
    float d0 = ds * (x1 * dy0 - y1 * dx0);
    float d1 = ds * (x2 * dy1 - y2 * dx1);
    float d2 = ds * (x0 * dy2 - y0 * dx2);
    if (ehb_same_sign(y1, y2)) { d0 = -EHB_F32_MAX; dy0 = 1.f; }
    if (ehb_same_sign(y2, y0)) { d1 = -EHB_F32_MAX; dy1 = 1.f; }
    if (ehb_same_sign(y0, y1)) { d2 = -EHB_F32_MAX; dy2 = 1.f; }
    const int di = ehb_max_idx3(d0, d1, d2, dy0, dy1, dy2);
    if (di == 0 && ehb_same_sign(a0, bb) && fabsf(dy0) >= fabsf(dx0)) dc = d0 / dy0;
    if (di == 1 && ehb_same_sign(a1, bb) && fabsf(dy1) >= fabsf(dx1)) dc = d1 / dy1;
    if (di == 2 && ehb_same_sign(a2, bb) && fabsf(dy2) >= fabsf(dx2)) dc = d2 / dy2;
    const float eps = .0625f;
    if (dc > -eps && dc < 1.f + eps) {
        dc = fminf(fmaxf(dc, 0.f), 1.f);
        *di_out = di;
        return ds * (.5f - dc);
    }
    return 0.f;
}

// Gradient of one pair's blend w.r.t. the clip positions (x, y, w) of the two vertices of its active edge.
// dd = dL/d(out of receiving pixel) * (c1 - c0).  Outputs the vertex indices and g1[3], g2[3] = (gx, gy, gw).
__device__ __noinline__ void ehb_aa_pair_grad(const EhbLink& lk, const float4* __restrict__ vc, int t, int side, int di,
                                              float al, float dd, int px, int py, int d, int H, int W, int* vi1_out,
                                              int* vi2_out, float* g1, float* g2)
{
    if (side) { px += 1 - d; py += d; }
    const int4 vi = __ldg(lk.faces + t);
    const int i1 = (di < 2) ? (di + 1) : 0, i2 = (i1 < 2) ? (i1 + 1) : 0;
    const int vi1 = i1 == 0 ? vi.x : (i1 == 1 ? vi.y : vi.z);
    const int vi2 = i2 == 0 ? vi.x : (i2 == 1 ? vi.y : vi.z);
    float p1v[4], p2v[4];
    ehb_clip_of(vc, vi1, p1v);
    ehb_clip_of(vc, vi2, p2v);
    float pxh = 0.5f * (float)W, pyh = 0.5f * (float)H;
    float fx = (float)px + .5f - pxh, fy = (float)py + .5f - pyh;
    if (d) {
        float s;
        s = p1v[0]; p1v[0] = p1v[1]; p1v[1] = s;
        s = p2v[0]; p2v[0] = p2v[1]; p2v[1] = s;
        s = pxh; pxh = pyh; pyh = s;
        s = fx; fx = fy; fy = s;
    }
    const float w1 = 1.f / p1v[3], w2 = 1.f / p2v[3];
    const float x1 = p1v[0] * w1 * pxh - fx, y1 = p1v[1] * w1 * pyh - fy;
    const float x2 = p2v[0] * w2 * pxh - fx, y2 = p2v[1] * w2 * pyh - fy;
    const float dx = x2 - x1, dy = y2 - y1;
    const float db = x1 * dy - y1 * dx;
    const float ep = copysignf(1e-3f, dy);
    const float iy = 1.f / (dy + ep);
    const float dby = db * iy;
    const float iw1 = -w1 * iy * dd, iw2 = w2 * iy * dd;
    float gp1x = iw1 * pxh * y2, gp2x = iw2 * pxh * y1;
    float gp1y = iw1 * pyh * (dby - x2), gp2y = iw2 * pyh * (dby - x1);
    float gp1w = -(p1v[0] * gp1x + p1v[1] * gp1y) * w1;
    float gp2w = -(p2v[0] * gp2x + p2v[1] * gp2y) * w2;
    if (d) { float s; s = gp1x; gp1x = gp1y; gp1y = s; s = gp2x; gp2x = gp2y; gp2y = s; }
    if (fabsf(al) >= 0.5f) { gp1x = gp1y = gp1w = 0.f; gp2x = gp2y = gp2w = 0.f; }
    *vi1_out = vi1; *vi2_out = vi2;
    g1[0] = gp1x; g1[1] = gp1y; g1[2] = gp1w;
    g2[0] = gp2x; g2[1] = gp2y; g2[2] = gp2w;
}
