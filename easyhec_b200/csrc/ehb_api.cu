// ehb_api.cu -- C ABI of libehb.so (see include/easyhec_b200.h for the contract of every entry point).
//
// Host side of the B200 silhouette rasterizer: context, mesh registry (padded float4/int4 buffers and the
// cached edge adjacency that replaces dr.antialias' per-call topology hash, nvdiffrast_renderer.py:43),
// scratch management and the launch sequence count -> alloc -> fill -> raster.  No torch types, no CPU
// fallback: every compute entry point enqueues CUDA kernels on the caller's stream or fails.
#include "../../include/easyhec_b200.h"
#include "ehb_kernels.cuh"
#include "ehb_tiles.cuh"
#include "ehb_pose.cuh"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                             \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) return fail(EHB_E_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct Ref {                           // reference masks registered once (ehb_ref_register)
    uint32_t* bits = nullptr;          // [B, H, ntx]  one bit per pixel, GL rows
    uint32_t* cnt = nullptr;           // [B, ntiles]
    unsigned long long* total = nullptr;   // [B]
    int B = 0, H = 0, W = 0, ntx = 0, ntiles = 0;
    bool live = false;
};

struct Mesh {
    float4* verts = nullptr;
    int4* faces = nullptr;
    int4* opp = nullptr;
    float4* boxes = nullptr;          // [2 * nboxes] AABBs of contiguous vertex chunks (device, recomputed with the vertices)
    float4* fboxes = nullptr;         // [2 * ceil(F/32)] AABBs of the 32-face batches (device, recomputed with the vertices)
    int V = 0, F = 0, nboxes = 0;
    bool live = false;
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    int ensure(size_t want, bool capturing)
    {
        if (want <= n) return EHB_OK;
        if (capturing) return fail(EHB_E_CAPACITY, "scratch too small during stream capture; call ehb_ctx_reserve first");
        if (p) CU(cudaFree(p));
        p = nullptr; n = 0;
        size_t cap = want + want / 4 + 256;
        CU(cudaMalloc((void**)&p, cap * sizeof(T)));
        n = cap;
        return EHB_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

constexpr double POOL_BUDGET = 8e9;  // bytes of plane pool per pipeline reserved without asking
constexpr int MAX_PIPES = 4;
#ifndef EHB_NSLOTS
#define EHB_NSLOTS 4
#endif
constexpr int N_SLOTS = EHB_NSLOTS;   // steps in flight (ehb_step_begin / ehb_solver_step_begin_* / _end)
constexpr int N_SCRATCH = MAX_PIPES + N_SLOTS;
constexpr size_t COMM_CH_WORDS = 2 * EHB_COMM_MAX * 8 + 8;   // one channel of a peer mailbox: [2 parities][ranks][8 words] + step counter

// Scratch of one pipeline.  A call's items are split over up to MAX_PIPES independent pipelines that run on internal
// streams: every kernel of the pass is latency / tail bound, so the pipelines fill each other's idle SMs.
struct Scratch {
    DevBuf<float4> vclip;             // pre-transformed vertices of one pass
    DevBuf<int2> vsnap;
    DevBuf<EhbPlane> plane;
    DevBuf<unsigned long long> pool;  // depth planes of one pass, bump-allocated
    DevBuf<uint32_t> tileList, emptyList, touch;
    DevBuf<EhbRec> bigRec;
    DevBuf<EhbUnit> units;
    DevBuf<uint32_t> batchBlk;        // parked heavy batches of k_raster
    DevBuf<uint32_t> batchList;       // visible batches of the pass (k_front -> k_raster)
    DevBuf<unsigned long long> bitPool;   // coverage bit per pixel of every plane (what the image-space stage reads)
    DevBuf<EhbBitsRec> bigBits;       // bit-plane address of every parked record
    DevBuf<unsigned char> pairPool;   // slabs for tiles whose silhouette pairs do not fit shared memory (k_tiles)
    EhbCounters* ctr = nullptr;
    void release()
    {
        vclip.release(); vsnap.release(); plane.release(); pool.release(); tileList.release(); emptyList.release();
        touch.release(); bigRec.release(); units.release(); batchBlk.release(); pairPool.release(); batchList.release(); bitPool.release(); bigBits.release();
    }
};

// Kinematic tree of a robot (ehb_robot_register): links in an order where a parent precedes its children.
#define EHB_KIN_MAX 64
struct KinTree {
    int n;
    int parent[EHB_KIN_MAX];      // -1: root
    int jtype[EHB_KIN_MAX];       // 0 fixed, 1 revolute / continuous, 2 prismatic
    int qidx[EHB_KIN_MAX];        // index into qpos of the joint that moves this link (-1: none)
    double mult[EHB_KIN_MAX], offs[EHB_KIN_MAX];   // joint value = qpos[qidx] * mult + offs (mimic joints)
    double axis[EHB_KIN_MAX][3];
    double origin[EHB_KIN_MAX][12];   // top three rows of the joint origin (parent frame -> joint frame)
};

struct Ctx {
    int device = 0;
    int nSM = 148;
    int occ = 4;                      // resident k_tiles CTAs per SM
    int occRaster = 4;                // resident k_raster CTAs per SM (persistent warps)
    int rule = 0;
    int nPipes = 3;
    bool pipesAuto = true;            // a call of <= 16 items runs as one pipeline unless the count was set explicitly (measured: r02)
    std::vector<Mesh> meshes;
    std::vector<Ref> refs;
    std::vector<KinTree*> robots;     // device copies of registered kinematic trees (ehb_robot_register)
    std::vector<int> robotLinks;
    DevBuf<double> fkScratch, fkIn;
    DevBuf<int> fkSel;
    Scratch sc[N_SCRATCH];            // [0, MAX_PIPES): pipelines of a device-pointer call; then one per host-step slot
    cudaStream_t pipeStream[MAX_PIPES] = {};
    cudaEvent_t evFork = nullptr, evJoin[MAX_PIPES] = {};
    EhbCounters* ctr = nullptr;       // device array [N_SCRATCH]
    EhbCounters* ctrHost = nullptr;   // pinned mirror [N_SCRATCH]
    cudaStream_t slotStream[N_SLOTS] = {};
    cudaEvent_t slotDone[N_SLOTS] = {};
    DevBuf<float> slotMvp[N_SLOTS];
    DevBuf<double> slotOut[N_SLOTS];
    DevBuf<uint8_t> slotRef[N_SLOTS];
    int slotB[N_SLOTS] = {}, slotL[N_SLOTS] = {};
    // staging for the host-buffer entry points
    DevBuf<float> mvpDev;
    DevBuf<double> outDev;            // loss[B] + gmvp[B*L*16]
    DevBuf<uint8_t> refDev, maskDev;
    DevBuf<unsigned long long> numDev;
    long long launches = 0;
    bool profiling = false;           // per-kernel CUDA events (ehb_ctx_profile): forces a single pipeline
    std::vector<cudaEvent_t> evPool;  // 5 events per profiled pass
    size_t evUsed = 0;
    double poolFactor = 2.0;          // beyond the budget: plane pool = items * H * W * poolFactor entries, grown on overflow
    double poolBudget = POOL_BUDGET;
    EhbComm comm = {};                // NVLink peer mailboxes (ehb_comm_*): channel 0, the caller's stream ...
    EhbComm commSlot[N_SLOTS] = {};   // ... and one channel per slot, so that the exchanges of steps in flight never mix
    unsigned int* commBox = nullptr;  // own mailbox (device)
    bool commReady = false;
    unsigned long long* dbgbuf = nullptr;   // EHB_TIMING builds
    unsigned int* hostFlags = nullptr;      // mapped pinned host words set by the kernels (ehb_ctx_poll)
    unsigned int* hostFlagsDev = nullptr;   // the same memory as the device sees it
};

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

bool is_capturing(cudaStream_t s)
{
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &st) != cudaSuccess) { cudaGetLastError(); return false; }
    return st != cudaStreamCaptureStatusNone;
}

// Edge adjacency: for triangle t and corner k, the vertex opposite to edge k (joining corners k+1, k+2) in the
// other triangle on that edge, or -1.  An edge keeps the first two opposite vertices in triangle order.
void build_adjacency(const int* tri, int F, int V, std::vector<int4>& opp)
{
    struct Rec { unsigned long long key; int order; int other; };
    std::vector<Rec> rec;
    rec.reserve((size_t)F * 3);
    for (int t = 0; t < F; t++) {
        const int v[3] = {tri[3 * t], tri[3 * t + 1], tri[3 * t + 2]};
        bool bad = false;
        for (int k = 0; k < 3; k++) bad |= (v[k] < 0 || v[k] >= V);
        if (bad || v[0] == v[1] || v[1] == v[2] || v[2] == v[0]) continue;
        for (int k = 0; k < 3; k++) {
            const int a = v[(k + 1) % 3], b = v[(k + 2) % 3];
            const unsigned long long lo = (unsigned)std::min(a, b), hi = (unsigned)std::max(a, b);
            rec.push_back({(lo << 32) | hi, 3 * t + k, v[k]});
        }
    }
    std::sort(rec.begin(), rec.end(), [](const Rec& x, const Rec& y) {
        return x.key != y.key ? x.key < y.key : x.order < y.order;
    });
    std::vector<int> flat((size_t)F * 3, -1);
    for (size_t i = 0; i < rec.size();) {
        size_t j = i;
        while (j < rec.size() && rec[j].key == rec[i].key) j++;
        const int n0 = rec[i].other, n1 = (j - i >= 2) ? rec[i + 1].other : -1;
        for (size_t k = i; k < j; k++) flat[rec[k].order] = (n0 == rec[k].other) ? n1 : n0;
        i = j;
    }
    opp.resize(F);
    for (int t = 0; t < F; t++) opp[t] = make_int4(flat[3 * t], flat[3 * t + 1], flat[3 * t + 2], 0);
}

__global__ void ehb_k_pad_verts(const float* __restrict__ src, float4* __restrict__ dst, int V)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < V) dst[i] = make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 1.f);
}

// AABB of each of nb contiguous vertex chunks (one warp per chunk) -> boxes[2*b] = min, boxes[2*b+1] = max
__global__ void ehb_k_boxes(const float4* __restrict__ verts, int V, int nb, float4* __restrict__ boxes)
{
    const int b = blockIdx.x, lane = threadIdx.x;
    const int cs = (V + nb - 1) / nb, i0 = b * cs, i1 = min(V, i0 + cs);
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int i = i0 + lane; i < i1; i += 32) {
        const float4 v = verts[i];
        lo[0] = fminf(lo[0], v.x); lo[1] = fminf(lo[1], v.y); lo[2] = fminf(lo[2], v.z);
        hi[0] = fmaxf(hi[0], v.x); hi[1] = fmaxf(hi[1], v.y); hi[2] = fmaxf(hi[2], v.z);
    }
    for (int o = 16; o > 0; o >>= 1)
        for (int k = 0; k < 3; k++) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    if (lane == 0) {
        if (i0 >= i1) { lo[0] = lo[1] = lo[2] = hi[0] = hi[1] = hi[2] = 0.f; }   // empty chunk: a point that is re-covered below
        boxes[2 * b] = make_float4(lo[0], lo[1], lo[2], 1.f);
        boxes[2 * b + 1] = make_float4(hi[0], hi[1], hi[2], 1.f);
    }
}

// AABB of each batch of 32 consecutive faces (one warp per batch, one face per lane)
__global__ void ehb_k_face_boxes(const float4* __restrict__ verts, int V, const int4* __restrict__ faces, int F,
                                 float4* __restrict__ fboxes)
{
    const int b = blockIdx.x, lane = threadIdx.x, f = b * 32 + lane;
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    if (f < F) {
        const int4 id = faces[f];
        const int vi[3] = {id.x, id.y, id.z};
        for (int k = 0; k < 3; k++)
            if ((unsigned)vi[k] < (unsigned)V) {   // faces with an invalid index are never drawn
                const float4 v = verts[vi[k]];
                lo[0] = fminf(lo[0], v.x); lo[1] = fminf(lo[1], v.y); lo[2] = fminf(lo[2], v.z);
                hi[0] = fmaxf(hi[0], v.x); hi[1] = fmaxf(hi[1], v.y); hi[2] = fmaxf(hi[2], v.z);
            }
    }
    for (int o = 16; o > 0; o >>= 1)
        for (int k = 0; k < 3; k++) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    if (lane == 0) {
        fboxes[2 * b] = make_float4(lo[0], lo[1], lo[2], 1.f);       // an empty batch keeps lo > hi: culled
        fboxes[2 * b + 1] = make_float4(hi[0], hi[1], hi[2], 1.f);
    }
}

// num[q] += sum over this block's pixels of k (C - k), k = number of cameras whose mask covers the pixel.
// sum_px unbiased_var_c(mask) = num / (C (C - 1)) exactly, so the reduction is integer and order-free.
__global__ void __launch_bounds__(256) ehb_k_variance(const uint8_t* __restrict__ masks, int C, long long n,
                                                      unsigned long long* __restrict__ num)
{
    const int q = blockIdx.y;
    const uint8_t* base = masks + (size_t)q * C * n;
    unsigned long long acc = 0;
    const bool vec = (n & 15) == 0 && (((uintptr_t)masks) & 15) == 0;
    if (vec) {
        const long long n16 = n >> 4;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
            unsigned k[16];
#pragma unroll
            for (int j = 0; j < 16; j++) k[j] = 0;
            for (int c = 0; c < C; c++) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + (size_t)c * n) + i);
                const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 16; j++) k[j] += ((w[j >> 2] >> (8 * (j & 3))) & 255u) != 0u;
            }
#pragma unroll
            for (int j = 0; j < 16; j++) acc += (unsigned long long)(k[j] * ((unsigned)C - k[j]));
        }
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
            unsigned k = 0;
            for (int c = 0; c < C; c++) k += __ldg(base + (size_t)c * n + i) != 0;
            acc += (unsigned long long)(k * ((unsigned)C - k));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(num + q, acc);
}

__global__ void ehb_k_variance_finish(const unsigned long long* __restrict__ num, int Q, int C, double* __restrict__ score)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < Q) score[q] = C > 1 ? (double)num[q] / ((double)C * (double)(C - 1)) : 0.0;
}


// Reference masks -> one bit per pixel (GL rows, a 32-bit word per (row, tile column)), set bits per tile interior and per
// item.  One warp per word, lane = pixel; image row r holds GL row H - 1 - r.  nonBinary counts f32 values outside {0, 1}.
template <typename T>
__global__ void ehb_k_pack_ref(const T* __restrict__ ref, int B, int H, int W, int ntx, uint32_t* __restrict__ bits,
                               uint32_t* __restrict__ cnt, unsigned long long* __restrict__ total,
                               unsigned long long* __restrict__ nonBinary)
{
    const long long word = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (word >= (long long)B * H * ntx) return;
    const int tx = (int)(word % ntx);
    const long long t = word / ntx;
    const int py = (int)(t % H), item = (int)(t / H);
    const int px = tx * EHB_T + lane;
    bool on = false, odd = false;
    if (px < W) {
        const T v = ref[((size_t)item * H + (H - 1 - py)) * W + px];
        on = v != (T)0;
        odd = on && v != (T)1;
    }
    const unsigned w = __ballot_sync(0xffffffffu, on), o = __ballot_sync(0xffffffffu, odd);
    if (lane == 0) {
        bits[word] = w;
        if (w) {
            atomicAdd(cnt + (size_t)item * ntx * ((H + EHB_T - 1) / EHB_T) + (size_t)(py / EHB_T) * ntx + tx, (unsigned)__popc(w));
            atomicAdd(total + item, (unsigned long long)__popc(w));
        }
        if (o && nonBinary && sizeof(T) == 4) atomicAdd(nonBinary, (unsigned long long)__popc(o));
    }
}

// ---------------------------------------------------------------------------------------------------- space exploration
__device__ __forceinline__ void ehb_mul34(const double* A, const double* B, double* C)   // C = A @ B, 3x4 rigid transforms
{
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 4; c++) {
            double s = A[4 * r] * B[c] + A[4 * r + 1] * B[4 + c] + A[4 * r + 2] * B[8 + c];
            if (c == 3) s += A[4 * r + 3];
            C[4 * r + c] = s;
        }
    }
}

// One thread per candidate joint configuration: forward kinematics of the whole tree in fp64 (URDF conventions of
// easyhec_b200/urdf_fk.py, which stands in for sapien / pinocchio: easyhec/structures/sapien_kin.py:26-30), then for every
// camera pose c and selected link l   mvp[q, c, l] = P @ cam[c] @ T_link   (nvdiffrast_renderer.py:33-37 with
// object_pose = Tc_c2b @ link_pose, render_api.py:179-190), rounded to fp32 once.
__global__ void ehb_k_fk_mvp(const KinTree* __restrict__ kt, const double* __restrict__ qpos, int dof, int Q,
                             const double* __restrict__ cams, int C, const double* __restrict__ P, const int* __restrict__ sel,
                             int L, float* __restrict__ mvp, double* __restrict__ scratch)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    double* T = scratch + (size_t)q * kt->n * 12;   // link poses of this candidate (global scratch: up to 64 x 12 doubles)
    for (int l = 0; l < kt->n; l++) {
        double J[12];
        const double* O = kt->origin[l];
        const int qi = kt->qidx[l];
        if (kt->jtype[l] != 0 && qi >= 0) {
            const double v = (qi < dof ? qpos[(size_t)q * dof + qi] : 0.0) * kt->mult[l] + kt->offs[l];
            double M[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
            const double x = kt->axis[l][0], y = kt->axis[l][1], z = kt->axis[l][2];
            if (kt->jtype[l] == 2) { M[3] = v * x; M[7] = v * y; M[11] = v * z; }
            else {   // Rodrigues: I + sin(v) K + (1 - cos(v)) K^2
                const double sn = sin(v), cs = 1.0 - cos(v);
                const double K[9] = {0, -z, y, z, 0, -x, -y, x, 0};
                for (int r = 0; r < 3; r++)
                    for (int c = 0; c < 3; c++) {
                        double k2 = 0.0;
                        for (int k = 0; k < 3; k++) k2 += K[3 * r + k] * K[3 * k + c];
                        M[4 * r + c] = (r == c ? 1.0 : 0.0) + sn * K[3 * r + c] + cs * k2;
                    }
            }
            ehb_mul34(O, M, J);
        } else {
            for (int i = 0; i < 12; i++) J[i] = O[i];
        }
        const int pa = kt->parent[l];
        if (pa < 0) { for (int i = 0; i < 12; i++) T[l * 12 + i] = J[i]; }
        else {
            double R[12];
            ehb_mul34(T + pa * 12, J, R);
            for (int i = 0; i < 12; i++) T[l * 12 + i] = R[i];
        }
    }
    for (int c = 0; c < C; c++)
        for (int k = 0; k < L; k++) {
            double CT[12];
            const double* cam = cams + (size_t)c * 16;
            ehb_mul34(cam, T + sel[k] * 12, CT);
            float* out = mvp + (((size_t)q * C + c) * L + k) * 16;
            for (int r = 0; r < 4; r++)
                for (int cc = 0; cc < 4; cc++) {
                    double s = P[4 * r] * CT[cc] + P[4 * r + 1] * CT[4 + cc] + P[4 * r + 2] * CT[8 + cc];
                    if (cc == 3) s += P[4 * r + 3];
                    out[4 * r + cc] = (float)s;
                }
        }
}

// Variance score straight from the depth planes of a packed-robot pass (one plane per (candidate, camera)): for every
// pixel of the union of the candidate's C plane boxes, k = number of cameras whose nearest triangle there has z/w > 0;
// num[q] += sum k (C - k).  Pixels outside every box have k = 0.  No mask is written anywhere.
__global__ void __launch_bounds__(256) ehb_k_variance_planes(const EhbPlane* __restrict__ plane, const unsigned long long* __restrict__ pool,
                                                             int C, unsigned long long* __restrict__ num)
{
    const int q = blockIdx.y;
    const EhbPlane* pl = plane + (size_t)q * C;
    int x0 = INT_MAX, y0 = INT_MAX, x1 = INT_MIN, y1 = INT_MIN;
    for (int c = 0; c < C; c++)
        if (pl[c].w > 0) {
            x0 = min(x0, pl[c].x0); y0 = min(y0, pl[c].y0);
            x1 = max(x1, pl[c].x0 + pl[c].w - 1); y1 = max(y1, pl[c].y0 + pl[c].h - 1);
        }
    unsigned long long acc = 0;
    if (x0 <= x1) {
        const int bw = x1 - x0 + 1;
        const long long n = (long long)bw * (y1 - y0 + 1);
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
            const int py = y0 + (int)(i / bw), px = x0 + (int)(i % bw);
            unsigned k = 0;
            for (int c = 0; c < C; c++) {
                const int cx = px - pl[c].x0, cy = py - pl[c].y0;
                if ((unsigned)cx < (unsigned)pl[c].w && (unsigned)cy < (unsigned)pl[c].h) {
                    const unsigned long long key = pool[pl[c].off + (long long)cy * pl[c].w + cx];
                    k += key != EHB_EMPTY && (uint32_t)(key >> 32) > 0x80000000u;
                }
            }
            acc += (unsigned long long)(k * ((unsigned)C - k));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(num + q, acc);
}

// Developer knob: integer from the environment, read once per name (launch-shape experiments without a rebuild).
int tune_int(const char* name, int dflt)
{
    static std::map<std::string, int> cache;
    auto it = cache.find(name);
    if (it != cache.end()) return it->second;
    const char* v = getenv(name);
    const int r = (v && *v) ? atoi(v) : dflt;
    cache[name] = r;
    return r;
}
constexpr int BIG_CAP = 1 << 17;     // deferred triangles per pass (16 MB of records)
constexpr int UNIT_CAP = 1 << 19;
constexpr int BATCH_CAP = 1 << 14;   // parked heavy batches per pass (70 MB); beyond it a batch is simply drawn inline

struct Io {
    const float* ref = nullptr; const uint8_t* ref_u8 = nullptr;
    float* masks = nullptr; double* loss = nullptr; double* gmvp = nullptr; float* gpos = nullptr;
    const float* dy = nullptr; uint8_t* out_u8 = nullptr; float* score = nullptr; int C = 0;
    int do_bwd = 0, clamp = 0; float invB = 1.f;
    int inflight = 0;                 // the pass is one of several in flight (slots): fewest instructions rather than shortest tails
    const uint32_t* refBits = nullptr; const uint32_t* refCnt = nullptr; const unsigned long long* refTotal = nullptr;
};

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
        else
            cudaGetLastError();
    }
    return fn;
}

// masks f32 [items][H][W] as a 3-D tensor with 32 x `rows` x 1 boxes.  false: not expressible (pitch / alignment) -> plain stores.
bool make_mask_map(CUtensorMap* map, float* masks, int items, int H, int W, int rows)
{
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || !masks || (W & 3) != 0 || (((uintptr_t)masks) & 15) != 0 || tune_int("EHB_NO_TMA", 0)) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)items};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * 4 * (cuuint64_t)H};
    const cuuint32_t box[3] = {EHB_T, (cuuint32_t)rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, masks, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Launch helper: kernels after the first of a pass are chained with programmatic dependent launch when EHB_PDL is on.
template <typename... KArgs, typename... Args>
cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool chained, Args... args)
{
#ifdef EHB_PDL
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = chained ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
#else
    (void)chained;
    kernel<<<grid, block, smem, st>>>(KArgs(args)...);
    return cudaGetLastError();
#endif
}

int build_robot(Ctx* c, const int* mesh_ids, int L, EhbRobot& rb)
{
    if (L < 1 || L > EHB_MAX_LINKS) return fail(EHB_E_ARG, "number of links %d outside [1, %d]", L, EHB_MAX_LINKS);
    memset(&rb, 0, sizeof rb);
    rb.L = L;
    for (int l = 0; l < L; l++) {
        const int id = mesh_ids[l];
        if (id < 0 || id >= (int)c->meshes.size() || !c->meshes[id].live) return fail(EHB_E_ARG, "unknown mesh id %d", id);
        const Mesh& m = c->meshes[id];
        if (m.F > (int)EHB_FACE_MASK) return fail(EHB_E_ARG, "mesh %d has too many faces", id);
        rb.link[l].verts = m.verts; rb.link[l].faces = m.faces; rb.link[l].opp = m.opp; rb.link[l].boxes = m.boxes; rb.link[l].nboxes = m.nboxes; rb.link[l].fboxes = m.fboxes;
        rb.link[l].V = m.V; rb.link[l].F = m.F;
        rb.foff[l + 1] = rb.foff[l] + m.F;
        rb.voff[l + 1] = rb.voff[l] + m.V;
        rb.boff[l + 1] = rb.boff[l] + (m.F + 31) / 32;
    }
    return EHB_OK;
}

int ensure_scratch(Ctx* c, Scratch& sc, int items, int L, int Lp, int H, int W, int Vtot, int Ftot, bool capturing)
{
    int r;
    if ((r = sc.vclip.ensure((size_t)items * std::max(Vtot, 1), capturing))) return r;
    if ((r = sc.vsnap.ensure((size_t)items * std::max(Vtot, 1), capturing))) return r;
    const int ntiles = ((W + EHB_T - 1) / EHB_T) * ((H + EHB_T - 1) / EHB_T);
    if (Lp == L) {   // image-space stage: pair pool (test mode, pool budget 0: one slab, so that the growth path runs)
        const size_t nSlabs = c->poolBudget == 0.0 ? (size_t)std::max(1.0, 16.0 * c->poolFactor)
                                                   : (size_t)(64.0 * std::max(1.0, c->poolFactor / 2.0));
        if ((r = sc.pairPool.ensure(nSlabs * EHB_SLAB_BYTES, capturing))) return r;
    }
    if ((r = sc.plane.ensure((size_t)items * Lp, capturing))) return r;
    if ((r = sc.tileList.ensure((size_t)items * ntiles, capturing))) return r;
    if ((r = sc.touch.ensure((size_t)items * ntiles, capturing))) return r;
    if ((r = sc.emptyList.ensure((size_t)items * ntiles, capturing))) return r;
    // queues of deferred triangles: a quarter of the pass's triangles may be parked, four units each on average
    // (test mode, pool budget 0: queues so small that the inline fallbacks of k_raster run -- slower, same result)
    const bool tinyQ = c->poolBudget == 0.0;
    // (capacities are per sub-queue, EHB_NQ of them: a warp's batches go to the sub-queue of its index, so they fill evenly)
    const size_t bigCap = tinyQ ? 1 : std::max<size_t>(BIG_CAP, std::min<size_t>((size_t)items * (size_t)std::max(Ftot, 1) / 4, (size_t)1 << 24)) / EHB_NQ;
    if ((r = sc.bigRec.ensure(bigCap * EHB_NQ, capturing))) return r;
    if ((r = sc.units.ensure((tinyQ ? 4 : std::max<size_t>(UNIT_CAP / EHB_NQ, 4 * bigCap)) * EHB_NQ, capturing))) return r;
    if ((r = sc.batchBlk.ensure((size_t)BATCH_CAP * EHB_BLK_WORDS, capturing))) return r;
    if ((r = sc.batchList.ensure((size_t)items * (size_t)(std::max(Ftot, 1) / 32 + EHB_MAX_LINKS + 1), capturing))) return r;
    // Plane pool: the worst case (every link's bbox is the whole screen) is items * Lp * H * W entries.  That is what is
    // reserved while it stays under POOL_BUDGET (180 GB of HBM: 10 views x 7 links x 1280x720 is 0.5 GB) -- then the
    // pool can never overflow; beyond it the pool holds poolFactor screens per item and grows on the overflow flag.
    const double worst = (double)items * Lp * H * W;
    const double f = worst * 8.0 <= c->poolBudget ? (double)Lp : std::min(std::max(c->poolFactor, c->poolBudget / (8.0 * items * H * W)), (double)Lp);
    // k_front clears only what the previous pass used: a new allocation starts clean (EMPTY is all ones) with zero extents
    const unsigned long long* pool0 = sc.pool.p;
    if ((r = sc.pool.ensure((size_t)((double)items * H * W * f) + 1024, capturing))) return r;
    if (sc.pool.p != pool0) {
        CU(cudaMemset(sc.pool.p, 0xFF, sc.pool.n * sizeof(unsigned long long)));
        CU(cudaMemset(&sc.ctr->prevCursor, 0, sizeof(unsigned long long)));
        CU(cudaDeviceSynchronize());
    }
    if (Lp == L) {   // coverage bits: a row of a plane is ceil(w / 64) <= w / 64 + 1 words
        const unsigned long long* bits0 = sc.bitPool.p;
        if ((r = sc.bitPool.ensure(sc.pool.n / 64 + (size_t)items * Lp * H + 1024, capturing))) return r;
        if (sc.bitPool.p != bits0) {
            CU(cudaMemset(sc.bitPool.p, 0, sc.bitPool.n * sizeof(unsigned long long)));
            CU(cudaMemset(&sc.ctr->prevBitCursor, 0, sizeof(unsigned long long)));
            CU(cudaDeviceSynchronize());
        }
        if ((r = sc.bigBits.ensure(sc.bigRec.n, capturing))) return r;
    }
    return EHB_OK;
}

int run_pass(Ctx* c, Scratch& sc, const int* mesh_ids, int L, int items, const float* mvp_dev, int H, int W, int mode,
             const Io& io, cudaStream_t st)
{
    if (!c) return fail(EHB_E_ARG, "null context");
    if (H < 1 || W < 1 || H > 8160 || W > 8160) return fail(EHB_E_ARG, "resolution %dx%d outside [1, 8160]", H, W);
    if (items < 0) return fail(EHB_E_ARG, "negative item count");
    if (items == 0) return EHB_OK;
    if (items > 65535) return fail(EHB_E_ARG, "more than 65535 items per launch");
    if (!mvp_dev) return fail(EHB_E_ARG, "null mvp pointer");
    DeviceGuard guard(c->device);
    EhbRobot rb;
    int r = build_robot(c, mesh_ids, L, rb);
    if (r) return r;
    const bool unionMode = mode == EHB_MODE_UNION;
    EhbParams p;
    memset(&p, 0, sizeof p);
    p.H = H; p.W = W;
    p.ntx = (W + EHB_T - 1) / EHB_T; p.nty = (H + EHB_T - 1) / EHB_T; p.ntiles = p.ntx * p.nty;
    p.items = items; p.L = L; p.Lp = unionMode ? 1 : L; p.Ftot = rb.foff[L]; p.Vtot = rb.voff[L];
    switch (mode) {
    case EHB_MODE_FUSED: p.hlo = 1; p.hhi = io.do_bwd ? 2 : 1; break;
    case EHB_MODE_AA_FWD: p.hlo = 1; p.hhi = 1; break;
    case EHB_MODE_AA_BWD: p.hlo = 0; p.hhi = 1; break;
    default: p.hlo = 0; p.hhi = 0; break;
    }
    p.xs = 2.f / (float)W; p.xo = 1.f / (float)W - 1.f; p.ys = 2.f / (float)H; p.yo = 1.f / (float)H - 1.f;
    p.mode = mode; p.rule = c->rule; p.do_bwd = io.do_bwd; p.clamp = io.clamp; p.invB = io.invB;
    const bool capturing = is_capturing(st);
    if ((r = ensure_scratch(c, sc, items, L, p.Lp, H, W, p.Vtot, p.Ftot, capturing))) return r;
    p.mvp = mvp_dev;
    p.vclip = sc.vclip.p; p.vsnap = sc.vsnap.p;
    p.plane = sc.plane.p; p.pool = sc.pool.p; p.poolCap = sc.pool.n;
    p.tileList = sc.tileList.p; p.emptyList = sc.emptyList.p; p.touch = unionMode ? nullptr : sc.touch.p; p.bigRec = sc.bigRec.p; p.units = sc.units.p; p.bigCap = (int)(sc.bigRec.n / EHB_NQ); p.unitCap = (int)(sc.units.n / EHB_NQ); p.batchBlk = tune_int("EHB_NO_OFFLOAD", 0) ? nullptr : sc.batchBlk.p; p.batchCap = c->poolBudget == 0.0 ? 1 : BATCH_CAP / EHB_NQ; p.ctr = sc.ctr; p.bits = unionMode ? nullptr : sc.bitPool.p; p.bitCap = unionMode ? 0 : sc.bitPool.n; p.bigBits = unionMode ? nullptr : sc.bigBits.p; p.batchList = sc.batchList.p; p.heavyArea = (float)tune_int("EHB_HEAVY_AREA", 1024);
    const int ovArea = tune_int("EHB_SMALL_AREA", -1), ovInline = tune_int("EHB_RINLINE", -1);   // (developer overrides)
    p.smallArea = ovArea >= 0 ? ovArea : (io.inflight ? EHB_SMALL_AREA_INFLIGHT : EHB_SMALL_AREA_SERIAL);
    p.inlineGroups = ovInline >= 1 ? ovInline : (io.inflight ? EHB_INLINE_INFLIGHT : EHB_INLINE_SERIAL);
    p.ref = io.ref; p.ref_u8 = io.ref_u8; p.masks = io.masks; p.loss = io.loss; p.gmvp = io.gmvp; p.gpos = io.gpos;
    p.dy = io.dy; p.out_u8 = io.out_u8;
    p.refBits = io.refBits; p.refCnt = io.refCnt; p.refTotal = io.refTotal;
    p.pairPool = sc.pairPool.p; p.nSlabs = (int)(sc.pairPool.n / EHB_SLAB_BYTES);
    p.useTma = (!unionMode && io.masks && make_mask_map(&p.tmMask, io.masks, items, H, W, EHB_T) &&
                ((H % EHB_T) == 0 || make_mask_map(&p.tmMaskTop, io.masks, items, H, W, H % EHB_T))) ? 1 : 0;
    p.dbgbuf = c->dbgbuf;
    p.hostFlags = c->hostFlagsDev;

    // Stage timing (ehb_ctx_profile): CUDA events around the stages on the launching stream.  Under stream capture they
    // become event-record nodes of the graph (cudaEventRecordExternal), so that a replayed pass is timed with the graph's
    // launch gaps instead of the eager launch overhead of four kernels and five event records.
    cudaEvent_t* ev = nullptr;
    if (c->profiling) {
        if (c->evUsed + 5 > c->evPool.size()) {
            if (capturing) return fail(EHB_E_CAPACITY, "event pool exhausted during stream capture");
            const size_t old = c->evPool.size();
            c->evPool.resize(old + 5 * 256);
            for (size_t i = old; i < c->evPool.size(); i++) CU(cudaEventCreate(&c->evPool[i]));
        }
        ev = &c->evPool[c->evUsed];
        c->evUsed += 5;
    }
    auto mark = [&](int i) {
        if (!ev) return;
        if (capturing) cudaEventRecordWithFlags(ev[i], st, cudaEventRecordExternal);
        else cudaEventRecord(ev[i], st);
    };
    const int chunks = std::max(1, rb.boff[L]);   // 32-triangle batches per item (a batch never straddles two links)
    // spare CTAs of the raster launch finish the tiles no link touches; with registered reference masks (or none) that is
    // only the zero fill of the masks -- nothing to do at all when no masks are wanted
    const bool legacyStream = mode == EHB_MODE_FUSED && (io.ref || io.ref_u8);
    p.fillEmpty = (!unionMode && mode != EHB_MODE_AA_BWD && !legacyStream && io.masks) ? 1 : 0;
    const int streamBlocks = legacyStream ? c->nSM * tune_int("EHB_STREAM_MULT", 2) : 0;
    const long long rasterBlocks = ((long long)chunks * items + EHB_RWARPS - 1) / EHB_RWARPS;
    mark(0); mark(1);
    const int tableBlocks = (items * p.Lp + 7) / 8;
    const int vchunks = std::max(1, (p.Vtot + 255) / 256);
    const int batchBlocks = (int)(((long long)chunks * items + 31) / 32);
    const int clearBlocks = c->nSM;
    const long long tileThreads = unionMode ? 0 : (long long)items * p.ntiles;   // one lane per tile
    const unsigned restBlocks = (unsigned)(vchunks * items + batchBlocks + clearBlocks + (tileThreads + 255) / 256);
    if (tune_int("EHB_FRONT_SPLIT", 0)) {   // developer switch: the table CTAs as a launch of their own
        CU(launch(ehb_k_front, dim3((unsigned)tableBlocks), dim3(256), 0, st, false, rb, p, tableBlocks, vchunks, batchBlocks, clearBlocks, chunks, 0));
        CU(launch(ehb_k_front, dim3(restBlocks), dim3(256), 0, st, false, rb, p, tableBlocks, vchunks, batchBlocks, clearBlocks, chunks, tableBlocks));
    } else {
        CU(launch(ehb_k_front, dim3((unsigned)tableBlocks + restBlocks), dim3(256), 0, st, false, rb, p, tableBlocks, vchunks, batchBlocks,
                  clearBlocks, chunks, 0));
    }
    mark(2);
    CU(launch(ehb_k_raster, dim3((unsigned)(streamBlocks + rasterBlocks)), dim3(EHB_RWARPS * 32), 0, st, true, rb, p, streamBlocks, chunks));
    CU(launch(ehb_k_raster_big, dim3(c->nSM * EHB_BMIN_BLOCKS), dim3(256), 0, st, true, p));
    mark(3);
    if (unionMode && !io.out_u8) {
        // planes only (space exploration scores them directly)
    } else if (unionMode) {
        const int nq = ((W + 3) / 4) * H;
        CU(launch(ehb_k_union_out, dim3((unsigned)std::min((nq + 255) / 256, 4 * c->nSM), (unsigned)items), dim3(256), 0, st, true, p));
    } else {
        const long long maxTiles = (long long)items * p.ntiles;
        const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(maxTiles, (long long)c->nSM * tune_int("EHB_TILE_CTAS", 14)));
        const int refKind = mode != EHB_MODE_FUSED ? 0 : (io.refBits ? 3 : (io.ref ? 1 : (io.ref_u8 ? 2 : 0)));
        // a small pass lasts as long as its heaviest tile: 256-thread CTAs halve that; a pass with thousands of listed tiles wants
        // the 128-thread CTAs (twice as many resident).  Measured: 10 views 640x480 78 -> 74 us, 10 views 1280x720 29 -> 34 us.
        if (maxTiles <= (long long)tune_int("EHB_TILE_WIDE", 4096))
            CU(launch(t256::ehb_tiles_kernel(mode, refKind, io.do_bwd != 0), dim3(grid), dim3(256), 0, st, true, rb, p));
        else
            CU(launch(t128::ehb_tiles_kernel(mode, refKind, io.do_bwd != 0), dim3(grid), dim3(128), 0, st, true, rb, p));
    }
    mark(4);
    c->launches += 4;   // front, raster, raster_big, tiles | union_out
    CU(cudaGetLastError());
    return EHB_OK;
}

// Split the items of one call over the context's pipelines (fork on internal streams, join back into `st`).
int run_split(Ctx* c, const int* mesh_ids, int L, int items, const float* mvp_dev, int H, int W, int mode, const Io& io,
              cudaStream_t st)
{
    if (!c) return fail(EHB_E_ARG, "null context");
    // few items: the forks / joins and the split tails cost more than the overlap gives (10 views: 94 us against 101 us)
    const int n = (c->profiling || items < 2 || (c->pipesAuto && items <= 16)) ? 1 : std::min(c->nPipes, items);
    if (n <= 1) return run_pass(c, c->sc[0], mesh_ids, L, items, mvp_dev, H, W, mode, io, st);
    DeviceGuard guard(c->device);
    const size_t px = (size_t)H * W;
    CU(cudaEventRecord(c->evFork, st));
    int first = 0;
    for (int k = 0; k < n; k++) {
        const int cnt = items / n + (k < items % n ? 1 : 0);
        Io s = io;
        if (s.ref) s.ref += first * px;
        if (s.ref_u8) s.ref_u8 += first * px;
        if (s.masks) s.masks += first * px;
        if (s.dy) s.dy += first * px;
        if (s.out_u8) s.out_u8 += first * px;
        if (s.loss) s.loss += first;
        if (s.gmvp) s.gmvp += (size_t)first * L * 16;
        if (s.refBits) {
            const int ntx = (W + EHB_T - 1) / EHB_T, nty = (H + EHB_T - 1) / EHB_T;
            s.refBits += (size_t)first * H * ntx; s.refCnt += (size_t)first * ntx * nty; s.refTotal += first;
        }
        cudaStream_t sk = k == 0 ? st : c->pipeStream[k];
        if (k > 0) CU(cudaStreamWaitEvent(sk, c->evFork, 0));
        const int r = run_pass(c, c->sc[k], mesh_ids, L, cnt, mvp_dev + (size_t)first * L * 16, H, W, mode, s, sk);
        if (r) return r;
        if (k > 0) {
            CU(cudaEventRecord(c->evJoin[k], sk));
            CU(cudaStreamWaitEvent(st, c->evJoin[k], 0));
        }
        first += cnt;
    }
    return EHB_OK;
}

}  // namespace

extern "C" {

int ehb_version(void) { return 100; }
const char* ehb_last_error(void) { return g_err.c_str(); }

int ehb_ctx_create(int device, ehb_ctx_t* out)
{
    if (!out) return fail(EHB_E_ARG, "null output pointer");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(EHB_E_CUDA, "no CUDA device available (this library has no CPU path)");
    }
    if (device < 0 || device >= n) return fail(EHB_E_ARG, "device %d out of range (%d devices)", device, n);
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(EHB_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    Ctx* c = new Ctx();
    c->device = device;
    c->nSM = prop.multiProcessorCount;
    if (getenv("EHB_PIPES")) c->pipesAuto = false;
    c->nPipes = std::max(1, std::min(MAX_PIPES, tune_int("EHB_PIPES", c->nPipes)));   // (developer switch; ehb_ctx_set_pipelines)
    CU(cudaMalloc((void**)&c->ctr, N_SCRATCH * sizeof(EhbCounters)));
    CU(cudaMemset(c->ctr, 0, N_SCRATCH * sizeof(EhbCounters)));
    CU(cudaMallocHost((void**)&c->ctrHost, N_SCRATCH * sizeof(EhbCounters)));
    if (cudaHostAlloc((void**)&c->hostFlags, 16 * sizeof(unsigned), cudaHostAllocMapped) == cudaSuccess &&
        cudaHostGetDevicePointer((void**)&c->hostFlagsDev, c->hostFlags, 0) == cudaSuccess) {
        for (int i = 0; i < 16; i++) c->hostFlags[i] = 0u;
    } else {
        cudaGetLastError();
        c->hostFlags = nullptr; c->hostFlagsDev = nullptr;
    }
    for (int k = 0; k < N_SCRATCH; k++) c->sc[k].ctr = c->ctr + k;
    for (int k = 0; k < N_SLOTS; k++) {
        CU(cudaStreamCreateWithFlags(&c->slotStream[k], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->slotDone[k], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming));
    for (int k = 0; k < MAX_PIPES; k++) {
        c->sc[k].ctr = c->ctr + k;
        CU(cudaStreamCreateWithFlags(&c->pipeStream[k], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&c->evJoin[k], cudaEventDisableTiming));
    }
    // the kernels of a pass run concurrently (pipelines) and want 130 - 210 KB of shared memory per SM for their resident
    // CTAs: ask for the largest carveout everywhere (left to the driver's heuristic, k_tiles ran one CTA per SM)
    for (int m = 0; m < 2; m++)
        for (int rk = 0; rk < 4; rk++)
            for (int b = 0; b < 2; b++)
            {
                CU(cudaFuncSetAttribute(t128::ehb_tiles_kernel(m ? EHB_MODE_AA_BWD : EHB_MODE_FUSED, rk, b != 0),
                                        cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
                CU(cudaFuncSetAttribute(t256::ehb_tiles_kernel(m ? EHB_MODE_AA_BWD : EHB_MODE_FUSED, rk, b != 0),
                                        cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
            }
    CU(cudaFuncSetAttribute(ehb_k_raster, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaFuncSetAttribute(ehb_k_raster_big, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occRaster, ehb_k_raster, EHB_RWARPS * 32, 0));
    c->occRaster = std::max(1, c->occRaster);
    *out = c;
    return EHB_OK;
}

int ehb_ctx_destroy(ehb_ctx_t h)
{
    Ctx* c = (Ctx*)h;
    if (!c) return EHB_OK;
    DeviceGuard guard(c->device);
    cudaDeviceSynchronize();
    for (auto& m : c->meshes) if (m.live) { cudaFree(m.verts); cudaFree(m.faces); cudaFree(m.opp); cudaFree(m.boxes); cudaFree(m.fboxes); }
    for (auto& r : c->refs) if (r.live) { cudaFree(r.bits); cudaFree(r.cnt); cudaFree(r.total); }
    for (auto k : c->robots) if (k) cudaFree(k);
    c->fkScratch.release(); c->fkIn.release(); c->fkSel.release();
    for (int k = 0; k < MAX_PIPES; k++) { cudaStreamDestroy(c->pipeStream[k]); cudaEventDestroy(c->evJoin[k]); }
    for (int k = 0; k < N_SCRATCH; k++) c->sc[k].release();
    for (int k = 0; k < N_SLOTS; k++) { cudaStreamDestroy(c->slotStream[k]); cudaEventDestroy(c->slotDone[k]); c->slotMvp[k].release(); c->slotOut[k].release(); c->slotRef[k].release(); }
    cudaEventDestroy(c->evFork);
    c->mvpDev.release(); c->outDev.release(); c->refDev.release(); c->maskDev.release(); c->numDev.release();
    for (auto e : c->evPool) cudaEventDestroy(e);
    cudaFree(c->ctr);
    cudaFreeHost(c->ctrHost);
    if (c->hostFlags) cudaFreeHost(c->hostFlags);
    delete c;
    return EHB_OK;
}

int ehb_ctx_set_fill_rule(ehb_ctx_t h, int rule)
{
    Ctx* c = (Ctx*)h;
    if (!c || (rule != 0 && rule != 1)) return fail(EHB_E_ARG, "bad context or rule");
    c->rule = rule;
    return EHB_OK;
}

int ehb_ctx_reserve(ehb_ctx_t h, int n_items, int n_links, int max_faces, int H, int W)
{
    Ctx* c = (Ctx*)h;
    if (!c || n_items < 1 || n_links < 1 || max_faces < 0 || H < 1 || W < 1) return fail(EHB_E_ARG, "bad reserve arguments");
    DeviceGuard guard(c->device);
    int r = 0;
    const int np = std::max(1, std::min(c->nPipes, n_items));
    for (int k = 0; k < np; k++)   // every pipeline gets its share of the items (and pipeline 0 the whole, for profiling runs)
        if ((r = ensure_scratch(c, c->sc[k], k == 0 ? n_items : (n_items + np - 1) / np, n_links, n_links, H, W,
                                std::max(1, max_faces), std::max(1, max_faces), false)))   // V <= 3 F
            return r;
    if ((r = c->mvpDev.ensure((size_t)n_items * n_links * 16, false))) return r;
    if ((r = c->outDev.ensure((size_t)n_items * (1 + n_links * 16), false))) return r;
    return EHB_OK;
}

int ehb_ctx_grow_scratch(ehb_ctx_t h)
{
    Ctx* c = (Ctx*)h;
    if (!c) return fail(EHB_E_ARG, "null context");
    c->poolFactor *= 2.0;
    return EHB_OK;
}

int ehb_ctx_debug_counters(ehb_ctx_t h, unsigned long long* out16, int reset)
{
    Ctx* c = (Ctx*)h;
    if (!c || !out16) return fail(EHB_E_ARG, "null pointer argument");
    DeviceGuard guard(c->device);
    CU(cudaDeviceSynchronize());
    EhbCounters hc;
    CU(cudaMemcpy(&hc, c->ctr, sizeof hc, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 16; i++) out16[i] = hc.dbg[i];
    if (reset) { for (int i = 0; i < 16; i++) hc.dbg[i] = 0; CU(cudaMemcpy(c->ctr, &hc, sizeof hc, cudaMemcpyHostToDevice)); }
    return EHB_OK;
}

int ehb_ctx_debug_buffer(ehb_ctx_t h, unsigned long long* out, int n_words)
{
    Ctx* c = (Ctx*)h;
    if (!c) return fail(EHB_E_ARG, "null context");
    DeviceGuard guard(c->device);
    const size_t words = (size_t)6 * EHB_TL_N * 2;   // EHB_TIMELINE builds: six regions of (start, end | smid) entries
    if (!c->dbgbuf) { CU(cudaMalloc((void**)&c->dbgbuf, words * 8)); CU(cudaMemset(c->dbgbuf, 0, words * 8)); return EHB_OK; }
    CU(cudaDeviceSynchronize());
    if (out && n_words > 0) CU(cudaMemcpy(out, c->dbgbuf, std::min((size_t)n_words, words) * 8, cudaMemcpyDeviceToHost));
    return EHB_OK;
}

int ehb_ctx_debug_marks(ehb_ctx_t h, unsigned* out16)
{
    Ctx* c = (Ctx*)h;
    if (!c || !out16 || !c->hostFlags) return fail(EHB_E_ARG, "no marks");
    for (int i = 0; i < 16; i++) out16[i] = ((volatile unsigned*)c->hostFlags)[i];
    return EHB_OK;
}

int ehb_ctx_profile(ehb_ctx_t h, int enable)
{
    Ctx* c = (Ctx*)h;
    if (!c) return fail(EHB_E_ARG, "null context");
    c->profiling = enable != 0;
    if (c->profiling && c->evPool.empty()) {   // (created here: a captured pass cannot create them)
        DeviceGuard guard(c->device);
        c->evPool.resize(5 * 256);
        for (size_t i = 0; i < c->evPool.size(); i++) CU(cudaEventCreate(&c->evPool[i]));
    }
    return EHB_OK;
}

static int kernel_times(Ctx* c, double* ms4, long long* n_passes, bool reset);

int ehb_ctx_kernel_times(ehb_ctx_t h, double* ms4, long long* n_passes) { return kernel_times((Ctx*)h, ms4, n_passes, true); }
/* the same without forgetting the recorded passes: for passes captured in a CUDA graph, whose events every replay records again */
int ehb_ctx_kernel_times_peek(ehb_ctx_t h, double* ms4, long long* n_passes) { return kernel_times((Ctx*)h, ms4, n_passes, false); }

static int kernel_times(Ctx* c, double* ms4, long long* n_passes, bool reset)
{
    if (!c || !ms4 || !n_passes) return fail(EHB_E_ARG, "null pointer argument");
    DeviceGuard guard(c->device);
    CU(cudaDeviceSynchronize());
    for (int k = 0; k < 4; k++) ms4[k] = 0.0;
    for (size_t i = 0; i + 5 <= c->evUsed; i += 5)
        for (int k = 0; k < 4; k++) {
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, c->evPool[i + k], c->evPool[i + k + 1]));
            ms4[k] += ms;
        }
    *n_passes = (long long)(c->evUsed / 5);
    if (reset) c->evUsed = 0;
    return EHB_OK;
}

int ehb_ctx_status(ehb_ctx_t h, unsigned* flags, long long* n_need_clip)
{
    Ctx* c = (Ctx*)h;
    if (!c) return fail(EHB_E_ARG, "null context");
    DeviceGuard guard(c->device);
    CU(cudaDeviceSynchronize());
    EhbCounters hc[N_SCRATCH];
    CU(cudaMemcpy(hc, c->ctr, sizeof hc, cudaMemcpyDeviceToHost));
    unsigned f = 0;
    long long nc = 0;
    for (int k = 0; k < N_SCRATCH; k++) { f |= hc[k].flags; nc += (long long)hc[k].nNeedClip; hc[k].flags = 0; hc[k].nNeedClip = 0; }
    if (flags) *flags = f;
    if (n_need_clip) *n_need_clip = nc;
    CU(cudaMemcpy(c->ctr, hc, sizeof hc, cudaMemcpyHostToDevice));
    if (c->hostFlags) for (int i = 0; i < 3; i++) c->hostFlags[i] = 0u;
    return EHB_OK;
}

int ehb_ctx_poll(ehb_ctx_t h, unsigned* flags)
{
    Ctx* c = (Ctx*)h;
    if (!c || !flags) return fail(EHB_E_ARG, "null pointer argument");
    unsigned f = 0;
    if (c->hostFlags)
        for (int i = 0; i < 3; i++)
            if (*reinterpret_cast<volatile unsigned int*>(c->hostFlags + i)) f |= 1u << i;
    *flags = f;
    return EHB_OK;
}

int ehb_ctx_set_pool_budget(ehb_ctx_t h, double bytes)
{
    Ctx* c = (Ctx*)h;
    if (!c || !(bytes >= 0)) return fail(EHB_E_ARG, "bad context or budget");
    c->poolBudget = bytes;
    c->poolFactor = bytes == 0.0 ? 1.0 / 16.0 : 2.0;   // budget 0: start from the minimum so that growth is exercised
    return EHB_OK;
}

int ehb_ctx_set_pipelines(ehb_ctx_t h, int n)
{
    Ctx* c = (Ctx*)h;
    if (!c || n < 1 || n > MAX_PIPES) return fail(EHB_E_ARG, "pipelines must be in [1, %d]", MAX_PIPES);
    c->nPipes = n;
    c->pipesAuto = false;
    return EHB_OK;
}

int ehb_mesh_register(ehb_ctx_t h, const float* verts, int V, const int* faces, int F, int* mesh_id)
{
    Ctx* c = (Ctx*)h;
    if (!c || !mesh_id || V < 0 || F < 0 || (V > 0 && !verts) || (F > 0 && !faces)) return fail(EHB_E_ARG, "bad mesh arguments");
    DeviceGuard guard(c->device);
    std::vector<float4> v4((size_t)std::max(V, 1));
    for (int i = 0; i < V; i++) v4[i] = make_float4(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2], 1.f);
    std::vector<int4> f4((size_t)std::max(F, 1));
    for (int i = 0; i < F; i++) f4[i] = make_int4(faces[3 * i], faces[3 * i + 1], faces[3 * i + 2], 0);
    std::vector<int4> opp;
    build_adjacency(faces, F, V, opp);
    opp.resize((size_t)std::max(F, 1));
    Mesh m;
    m.V = V; m.F = F;
    CU(cudaMalloc((void**)&m.verts, v4.size() * sizeof(float4)));
    CU(cudaMalloc((void**)&m.faces, f4.size() * sizeof(int4)));
    CU(cudaMalloc((void**)&m.opp, opp.size() * sizeof(int4)));
    CU(cudaMemcpy(m.verts, v4.data(), v4.size() * sizeof(float4), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(m.faces, f4.data(), f4.size() * sizeof(int4), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(m.opp, opp.data(), opp.size() * sizeof(int4), cudaMemcpyHostToDevice));
    m.nboxes = V > 0 ? std::max(1, std::min(32, (V + 63) / 64)) : 0;
    CU(cudaMalloc((void**)&m.boxes, 2 * 32 * sizeof(float4)));
    CU(cudaMalloc((void**)&m.fboxes, 2 * (size_t)std::max(1, (F + 31) / 32) * sizeof(float4)));
    if (m.nboxes > 0) ehb_k_boxes<<<m.nboxes, 32>>>(m.verts, V, m.nboxes, m.boxes);
    if (F > 0) ehb_k_face_boxes<<<(F + 31) / 32, 32>>>(m.verts, V, m.faces, F, m.fboxes);
    CU(cudaDeviceSynchronize());
    m.live = true;
    int id = -1;
    for (size_t i = 0; i < c->meshes.size(); i++) if (!c->meshes[i].live) { id = (int)i; break; }
    if (id < 0) { id = (int)c->meshes.size(); c->meshes.push_back(m); } else c->meshes[id] = m;
    *mesh_id = id;
    return EHB_OK;
}

int ehb_mesh_update_verts(ehb_ctx_t h, int mesh_id, const float* verts_dev, int V, void* stream)
{
    Ctx* c = (Ctx*)h;
    if (!c || mesh_id < 0 || mesh_id >= (int)c->meshes.size() || !c->meshes[mesh_id].live) return fail(EHB_E_ARG, "unknown mesh id %d", mesh_id);
    Mesh& m = c->meshes[mesh_id];
    if (V != m.V || (V > 0 && !verts_dev)) return fail(EHB_E_ARG, "vertex count %d does not match the registered mesh (%d)", V, m.V);
    if (V == 0) return EHB_OK;
    DeviceGuard guard(c->device);
    ehb_k_pad_verts<<<(V + 255) / 256, 256, 0, (cudaStream_t)stream>>>(verts_dev, m.verts, V);
    ehb_k_boxes<<<m.nboxes, 32, 0, (cudaStream_t)stream>>>(m.verts, V, m.nboxes, m.boxes);
    if (m.F > 0) ehb_k_face_boxes<<<(m.F + 31) / 32, 32, 0, (cudaStream_t)stream>>>(m.verts, V, m.faces, m.F, m.fboxes);
    c->launches += 3;
    CU(cudaGetLastError());
    return EHB_OK;
}

int ehb_mesh_release(ehb_ctx_t h, int mesh_id)
{
    Ctx* c = (Ctx*)h;
    if (!c || mesh_id < 0 || mesh_id >= (int)c->meshes.size() || !c->meshes[mesh_id].live) return fail(EHB_E_ARG, "unknown mesh id %d", mesh_id);
    DeviceGuard guard(c->device);
    CU(cudaDeviceSynchronize());
    Mesh& m = c->meshes[mesh_id];
    cudaFree(m.verts); cudaFree(m.faces); cudaFree(m.opp); cudaFree(m.boxes); cudaFree(m.fboxes);
    m = Mesh();
    return EHB_OK;
}

int ehb_mesh_info(ehb_ctx_t h, int mesh_id, int* V, int* F)
{
    Ctx* c = (Ctx*)h;
    if (!c || mesh_id < 0 || mesh_id >= (int)c->meshes.size() || !c->meshes[mesh_id].live) return fail(EHB_E_ARG, "unknown mesh id %d", mesh_id);
    if (V) *V = c->meshes[mesh_id].V;
    if (F) *F = c->meshes[mesh_id].F;
    return EHB_OK;
}

int ehb_render_mask_fwd(ehb_ctx_t h, int mesh_id, const float* mvp_dev, int H, int W, int anti_aliasing, void* out_dev,
                        void* stream)
{
    if (!out_dev) return fail(EHB_E_ARG, "null output pointer");
    Io io;
    if (anti_aliasing) { io.masks = (float*)out_dev; return run_split((Ctx*)h, &mesh_id, 1, 1, mvp_dev, H, W, EHB_MODE_AA_FWD, io, (cudaStream_t)stream); }
    io.out_u8 = (uint8_t*)out_dev;
    return run_split((Ctx*)h, &mesh_id, 1, 1, mvp_dev, H, W, EHB_MODE_UNION, io, (cudaStream_t)stream);
}

int ehb_render_mask_bwd(ehb_ctx_t h, int mesh_id, const float* mvp_dev, int H, int W, const float* dy_dev,
                        double* g_mvp_dev, float* g_pos_dev, void* stream)
{
    Ctx* c = (Ctx*)h;
    if (!c || !dy_dev || !g_mvp_dev) return fail(EHB_E_ARG, "null pointer argument");
    if (g_pos_dev) {
        int V = 0;
        int r = ehb_mesh_info(h, mesh_id, &V, nullptr);
        if (r) return r;
        DeviceGuard guard(c->device);
        if (V > 0) CU(cudaMemsetAsync(g_pos_dev, 0, (size_t)V * 4 * sizeof(float), (cudaStream_t)stream));
    }
    Io io;
    io.dy = dy_dev; io.gmvp = g_mvp_dev; io.gpos = g_pos_dev; io.do_bwd = 1;
    return run_split(c, &mesh_id, 1, 1, mvp_dev, H, W, EHB_MODE_AA_BWD, io, (cudaStream_t)stream);
}

int ehb_render_views_fused(ehb_ctx_t h, const int* mesh_ids, int L, int B, const float* mvp_dev, const float* ref_dev,
                           int H, int W, int do_bwd, float* masks_dev, double* loss_dev, double* g_mvp_dev, void* stream)
{
    if (!mesh_ids) return fail(EHB_E_ARG, "null mesh id list");
    if (do_bwd && (!ref_dev || !g_mvp_dev)) return fail(EHB_E_ARG, "backward needs ref_dev and g_mvp_dev");
    if (ref_dev && !loss_dev) return fail(EHB_E_ARG, "ref_dev given without loss_dev");
    Io io;
    io.ref = ref_dev; io.masks = masks_dev; io.loss = ref_dev ? loss_dev : nullptr; io.gmvp = do_bwd ? g_mvp_dev : nullptr;
    io.do_bwd = do_bwd ? 1 : 0; io.clamp = 1; io.invB = B > 0 ? 1.0f / (float)B : 1.f;
    return run_split((Ctx*)h, mesh_ids, L, B, mvp_dev, H, W, EHB_MODE_FUSED, io, (cudaStream_t)stream);
}

int ehb_render_views_fused_u8(ehb_ctx_t h, const int* mesh_ids, int L, int B, const float* mvp_dev,
                              const uint8_t* ref_u8_dev, int H, int W, int do_bwd, float* masks_dev, double* loss_dev,
                              double* g_mvp_dev, void* stream)
{
    if (!mesh_ids) return fail(EHB_E_ARG, "null mesh id list");
    if (!ref_u8_dev || !loss_dev || (do_bwd && !g_mvp_dev)) return fail(EHB_E_ARG, "null pointer argument");
    Io io;
    io.ref_u8 = ref_u8_dev; io.masks = masks_dev; io.loss = loss_dev; io.gmvp = do_bwd ? g_mvp_dev : nullptr;
    io.do_bwd = do_bwd ? 1 : 0; io.clamp = 1; io.invB = B > 0 ? 1.0f / (float)B : 1.f;
    return run_split((Ctx*)h, mesh_ids, L, B, mvp_dev, H, W, EHB_MODE_FUSED, io, (cudaStream_t)stream);
}


int ehb_ref_register(ehb_ctx_t h, const void* ref, int dtype, int on_device, int B, int H, int W, int* ref_id)
{
    Ctx* c = (Ctx*)h;
    if (!c || !ref || !ref_id || B < 1 || H < 1 || W < 1 || H > 8160 || W > 8160 || (dtype != EHB_REF_U8 && dtype != EHB_REF_F32))
        return fail(EHB_E_ARG, "bad reference-mask arguments");
    DeviceGuard guard(c->device);
    const size_t esz = dtype == EHB_REF_F32 ? 4 : 1, n = (size_t)B * H * W;
    void* tmp = nullptr;
    const void* src = ref;
    if (!on_device) {
        CU(cudaMalloc(&tmp, n * esz));
        if (cudaMemcpy(tmp, ref, n * esz, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(tmp); return fail(EHB_E_CUDA, "copy of the reference masks failed"); }
        src = tmp;
    }
    Ref r;
    r.B = B; r.H = H; r.W = W; r.ntx = (W + EHB_T - 1) / EHB_T; r.ntiles = r.ntx * ((H + EHB_T - 1) / EHB_T);
    const size_t nw = (size_t)B * H * r.ntx;
    unsigned long long* bad = nullptr;
    CU(cudaMalloc((void**)&r.bits, nw * 4));
    CU(cudaMalloc((void**)&r.cnt, (size_t)B * r.ntiles * 4));
    CU(cudaMalloc((void**)&r.total, (size_t)B * 8));
    CU(cudaMalloc((void**)&bad, 8));
    CU(cudaMemset(r.cnt, 0, (size_t)B * r.ntiles * 4));
    CU(cudaMemset(r.total, 0, (size_t)B * 8));
    CU(cudaMemset(bad, 0, 8));
    const unsigned blocks = (unsigned)((nw * 32 + 255) / 256);
    if (dtype == EHB_REF_F32) ehb_k_pack_ref<float><<<blocks, 256>>>((const float*)src, B, H, W, r.ntx, r.bits, r.cnt, r.total, bad);
    else ehb_k_pack_ref<uint8_t><<<blocks, 256>>>((const uint8_t*)src, B, H, W, r.ntx, r.bits, r.cnt, r.total, bad);
    unsigned long long nbad = 0;
    cudaError_t e = cudaMemcpy(&nbad, bad, 8, cudaMemcpyDeviceToHost);
    cudaFree(bad);
    if (tmp) cudaFree(tmp);
    c->launches += 1;
    if (e != cudaSuccess || nbad) {
        cudaFree(r.bits); cudaFree(r.cnt); cudaFree(r.total);
        if (e != cudaSuccess) return fail(EHB_E_CUDA, "packing the reference masks failed: %s", cudaGetErrorString(e));
        return fail(EHB_E_ARG, "%llu reference-mask values are neither 0 nor 1: only binary masks can be registered "
                               "(use ehb_render_views_fused for soft references)", nbad);
    }
    r.live = true;
    int id = -1;
    for (size_t i = 0; i < c->refs.size(); i++) if (!c->refs[i].live) { id = (int)i; break; }
    if (id < 0) { id = (int)c->refs.size(); c->refs.push_back(r); } else c->refs[id] = r;
    *ref_id = id;
    return EHB_OK;
}

int ehb_ref_release(ehb_ctx_t h, int ref_id)
{
    Ctx* c = (Ctx*)h;
    if (!c || ref_id < 0 || ref_id >= (int)c->refs.size() || !c->refs[ref_id].live) return fail(EHB_E_ARG, "unknown reference id %d", ref_id);
    DeviceGuard guard(c->device);
    CU(cudaDeviceSynchronize());
    Ref& r = c->refs[ref_id];
    cudaFree(r.bits); cudaFree(r.cnt); cudaFree(r.total);
    r = Ref();
    return EHB_OK;
}

static int ref_io(Ctx* c, int ref_id, int first, int B, int H, int W, Io& io)
{
    if (!c || ref_id < 0 || ref_id >= (int)c->refs.size() || !c->refs[ref_id].live) return fail(EHB_E_ARG, "unknown reference id %d", ref_id);
    const Ref& r = c->refs[ref_id];
    if (r.H != H || r.W != W) return fail(EHB_E_ARG, "reference masks are %dx%d, the call renders %dx%d", r.H, r.W, H, W);
    if (first < 0 || B < 0 || first + B > r.B) return fail(EHB_E_ARG, "views [%d, %d) outside the %d registered reference masks", first, first + B, r.B);
    io.refBits = r.bits + (size_t)first * H * r.ntx;
    io.refCnt = r.cnt + (size_t)first * r.ntiles;
    io.refTotal = r.total + first;
    return EHB_OK;
}

int ehb_render_views_fused_ref(ehb_ctx_t h, const int* mesh_ids, int L, int B, const float* mvp_dev, int ref_id, int first_view,
                               int H, int W, int do_bwd, float* masks_dev, double* loss_dev, double* g_mvp_dev, void* stream)
{
    if (!mesh_ids) return fail(EHB_E_ARG, "null mesh id list");
    if (!loss_dev || (do_bwd && !g_mvp_dev)) return fail(EHB_E_ARG, "null pointer argument");
    Io io;
    int r = ref_io((Ctx*)h, ref_id, first_view, B, H, W, io);
    if (r) return r;
    io.masks = masks_dev; io.loss = loss_dev; io.gmvp = do_bwd ? g_mvp_dev : nullptr;
    io.do_bwd = do_bwd ? 1 : 0; io.clamp = 1; io.invB = B > 0 ? 1.0f / (float)B : 1.f;
    return run_split((Ctx*)h, mesh_ids, L, B, mvp_dev, H, W, EHB_MODE_FUSED, io, (cudaStream_t)stream);
}

int ehb_render_binary_batch(ehb_ctx_t h, const int* mesh_ids, int L, int N, const float* mvp_dev, int H, int W,
                            uint8_t* out_dev, void* stream)
{
    if (!mesh_ids || !out_dev) return fail(EHB_E_ARG, "null pointer argument");
    Io io;
    io.out_u8 = out_dev;
    return run_split((Ctx*)h, mesh_ids, L, N, mvp_dev, H, W, EHB_MODE_UNION, io, (cudaStream_t)stream);
}

int ehb_variance_score(ehb_ctx_t h, const uint8_t* masks_dev, int Q, int C, long long n, double* score_dev, void* stream)
{
    Ctx* c = (Ctx*)h;
    if (!c || Q < 0 || C < 1 || n < 0 || (Q > 0 && (!masks_dev || !score_dev))) return fail(EHB_E_ARG, "bad variance arguments");
    if (Q == 0) return EHB_OK;
    if (Q > 65535) return fail(EHB_E_ARG, "more than 65535 candidates per call");
    DeviceGuard guard(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    int r = c->numDev.ensure((size_t)Q, is_capturing(st));
    if (r) return r;
    CU(cudaMemsetAsync(c->numDev.p, 0, (size_t)Q * sizeof(unsigned long long), st));
    const long long work = (n + 15) / 16;
    const unsigned gx = (unsigned)std::max<long long>(1, std::min<long long>((work + 255) / 256, 4 * c->nSM));
    ehb_k_variance<<<dim3(gx, (unsigned)Q), 256, 0, st>>>(masks_dev, C, n, c->numDev.p);
    ehb_k_variance_finish<<<(Q + 255) / 256, 256, 0, st>>>(c->numDev.p, Q, C, score_dev);
    c->launches += 2;
    CU(cudaGetLastError());
    return EHB_OK;
}

int ehb_explore_scores(ehb_ctx_t h, const int* mesh_ids, int L, int Q, int C, const float* mvp_dev, int H, int W,
                       double* score_dev, void* stream)
{
    Ctx* c = (Ctx*)h;
    if (!c || !mesh_ids || Q < 0 || C < 1 || (Q > 0 && (!mvp_dev || !score_dev))) return fail(EHB_E_ARG, "bad explore arguments");
    if (Q == 0) return EHB_OK;
    DeviceGuard guard(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    const bool capturing = is_capturing(st);
    int r = c->numDev.ensure((size_t)Q, capturing);
    if (r) return r;
    CU(cudaMemsetAsync(c->numDev.p, 0, (size_t)Q * sizeof(unsigned long long), st));
    // Candidates are rendered in chunks (bounded scratch: one depth plane per (candidate, camera), no mask anywhere); the
    // chunks alternate over the context's pipelines, each on its own stream, and every chunk is scored straight from its
    // planes by ehb_k_variance_planes on that stream.
    const int chunk = std::max(1, std::min(Q, std::max(1, 64 / C)));
    const int np = c->profiling ? 1 : std::max(1, std::min(c->nPipes, (Q + chunk - 1) / chunk));
    CU(cudaEventRecord(c->evFork, st));
    for (int k = 1; k < np; k++) CU(cudaStreamWaitEvent(c->pipeStream[k], c->evFork, 0));
    int turn = 0;
    for (int q0 = 0; q0 < Q; q0 += chunk, turn++) {
        const int nq = std::min(chunk, Q - q0), k = turn % np;
        cudaStream_t sk = k == 0 ? st : c->pipeStream[k];
        Io io;   // union mode without an output image: planes only
        r = run_pass(c, c->sc[k], mesh_ids, L, nq * C, mvp_dev + (size_t)q0 * C * L * 16, H, W, EHB_MODE_UNION, io, sk);
        if (r) return r;
        ehb_k_variance_planes<<<dim3(32, (unsigned)nq), 256, 0, sk>>>(c->sc[k].plane.p, c->sc[k].pool.p, C, c->numDev.p + q0);
        c->launches += 1;
    }
    for (int k = 1; k < np; k++) {
        CU(cudaEventRecord(c->evJoin[k], c->pipeStream[k]));
        CU(cudaStreamWaitEvent(st, c->evJoin[k], 0));
    }
    ehb_k_variance_finish<<<(Q + 255) / 256, 256, 0, st>>>(c->numDev.p, Q, C, score_dev);
    c->launches += 1;
    CU(cudaGetLastError());
    return EHB_OK;
}

int ehb_robot_register(ehb_ctx_t h, int n_links, const int* parent, const int* jtype, const int* qidx, const double* mult,
                       const double* offs, const double* axis, const double* origin, int* robot_id)
{
    Ctx* c = (Ctx*)h;
    if (!c || !robot_id || n_links < 1 || n_links > EHB_KIN_MAX || !parent || !jtype || !qidx || !mult || !offs || !axis || !origin)
        return fail(EHB_E_ARG, "bad robot arguments (1 <= links <= %d)", EHB_KIN_MAX);
    KinTree kt;
    memset(&kt, 0, sizeof kt);
    kt.n = n_links;
    for (int l = 0; l < n_links; l++) {
        if (parent[l] >= l || parent[l] < -1) return fail(EHB_E_ARG, "link %d: parents must precede their children", l);
        kt.parent[l] = parent[l]; kt.jtype[l] = jtype[l]; kt.qidx[l] = qidx[l]; kt.mult[l] = mult[l]; kt.offs[l] = offs[l];
        for (int k = 0; k < 3; k++) kt.axis[l][k] = axis[3 * l + k];
        for (int k = 0; k < 12; k++) kt.origin[l][k] = origin[16 * l + k];
    }
    DeviceGuard guard(c->device);
    KinTree* d = nullptr;
    CU(cudaMalloc((void**)&d, sizeof kt));
    CU(cudaMemcpy(d, &kt, sizeof kt, cudaMemcpyHostToDevice));
    c->robots.push_back(d);
    c->robotLinks.push_back(n_links);
    *robot_id = (int)c->robots.size() - 1;
    return EHB_OK;
}

int ehb_explore_fk_mvp(ehb_ctx_t h, int robot_id, const double* qpos_dev, int dof, int Q, const double* cams_host, int C,
                       const float* K_host, int H, int W, const int* sel_links, int L, float* mvp_dev, void* stream)
{
    Ctx* c = (Ctx*)h;
    if (!c || robot_id < 0 || robot_id >= (int)c->robots.size() || !qpos_dev || !cams_host || !K_host || !sel_links || !mvp_dev ||
        Q < 0 || C < 1 || L < 1 || dof < 0)
        return fail(EHB_E_ARG, "bad explore_fk_mvp arguments");
    if (Q == 0) return EHB_OK;
    const int n = c->robotLinks[robot_id];
    for (int k = 0; k < L; k++)
        if (sel_links[k] < 0 || sel_links[k] >= n) return fail(EHB_E_ARG, "selected link %d outside the robot's %d links", sel_links[k], n);
    DeviceGuard guard(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    int r;
    if ((r = c->fkScratch.ensure((size_t)Q * n * 12, false))) return r;
    if ((r = c->fkIn.ensure((size_t)C * 16 + 16, false))) return r;
    if ((r = c->fkSel.ensure((size_t)L, false))) return r;
    // P = K_to_projection(K, H, W, n = 0.001, f = 10) @ diag(1, -1, -1, 1), as fp32 arithmetic like the host composes it
    // (easyhec/utils/nvdiffrast_utils.py:5-11), then widened
    std::vector<double> in((size_t)C * 16 + 16);
    {
        const float fu = K_host[0], fv = K_host[4], cu = K_host[2], cv = K_host[5];
        const float a = (float)(-(10.0 + 0.001) / (10.0 - 0.001)), b = (float)(-2.0 * 10.0 * 0.001 / (10.0 - 0.001));
        double* P = in.data() + (size_t)C * 16;
        for (int i = 0; i < 16; i++) P[i] = 0.0;
        P[0] = (double)((2.f * fu) / (float)W);
        P[2] = (double)(-((-2.f * cu) / (float)W + 1.f));
        P[5] = (double)(-((2.f * fv) / (float)H));
        P[6] = (double)(-((2.f * cv) / (float)H - 1.f));
        P[10] = (double)(-a);
        P[11] = (double)b;
        P[14] = 1.0;
    }
    memcpy(in.data(), cams_host, (size_t)C * 16 * sizeof(double));
    CU(cudaMemcpyAsync(c->fkIn.p, in.data(), in.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->fkSel.p, sel_links, (size_t)L * sizeof(int), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));   // (the staging vectors are pageable host memory)
    ehb_k_fk_mvp<<<(Q + 63) / 64, 64, 0, st>>>(c->robots[robot_id], qpos_dev, dof, Q, c->fkIn.p, C, c->fkIn.p + (size_t)C * 16,
                                                c->fkSel.p, L, mvp_dev, c->fkScratch.p);
    c->launches += 1;
    CU(cudaGetLastError());
    return EHB_OK;
}

int ehb_solver_step_host(ehb_ctx_t h, const int* mesh_ids, int L, int B, const float* mvp_host, const float* ref_dev,
                         int H, int W, double* loss_host, double* g_mvp_host, void* stream)
{
    Ctx* c = (Ctx*)h;
    if (!c || !mvp_host || !ref_dev || !loss_host || !g_mvp_host) return fail(EHB_E_ARG, "null pointer argument");
    if (B < 1 || L < 1) return fail(EHB_E_ARG, "empty batch");
    DeviceGuard guard(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    int r;
    const size_t nm = (size_t)B * L * 16, no = (size_t)B + nm;
    if ((r = c->mvpDev.ensure(nm, false))) return r;
    if ((r = c->outDev.ensure(no, false))) return r;
    CU(cudaMemcpyAsync(c->mvpDev.p, mvp_host, nm * sizeof(float), cudaMemcpyHostToDevice, st));
    for (int attempt = 0;; attempt++) {
        r = ehb_render_views_fused(h, mesh_ids, L, B, c->mvpDev.p, ref_dev, H, W, 1, nullptr, c->outDev.p, c->outDev.p + B, st);
        if (r) return r;
        CU(cudaMemcpyAsync(loss_host, c->outDev.p, B * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(g_mvp_host, c->outDev.p + B, nm * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(c->ctrHost, c->ctr, MAX_PIPES * sizeof(EhbCounters), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        { unsigned f = 0; for (int k = 0; k < MAX_PIPES; k++) f |= c->ctrHost[k].flags; if (!(f & EHB_FLAG_POOL_OVERFLOW)) break; }
        if (attempt >= 6) return fail(EHB_E_OVERFLOW, "plane pool overflow persists after growing");
        c->poolFactor *= 2.0;   // grow the plane pool and run the step again
        for (int k = 0; k < MAX_PIPES; k++) CU(cudaMemsetAsync(&c->ctr[k].flags, 0, sizeof(unsigned), st));
    }
    return EHB_OK;
}

int ehb_solver_step_host_u8(ehb_ctx_t h, const int* mesh_ids, int L, int B, const float* mvp_host,
                             const uint8_t* ref_u8_host, int H, int W, double* loss_host, double* g_mvp_host, void* stream)
{
    Ctx* c = (Ctx*)h;
    if (!c || !mvp_host || !ref_u8_host || !loss_host || !g_mvp_host) return fail(EHB_E_ARG, "null pointer argument");
    if (B < 1 || L < 1) return fail(EHB_E_ARG, "empty batch");
    DeviceGuard guard(c->device);
    cudaStream_t st = (cudaStream_t)stream;
    int r;
    const size_t nm = (size_t)B * L * 16, no = (size_t)B + nm, npx = (size_t)B * H * W;
    if ((r = c->mvpDev.ensure(nm, false))) return r;
    if ((r = c->outDev.ensure(no, false))) return r;
    if ((r = c->refDev.ensure(npx, false))) return r;
    CU(cudaMemcpyAsync(c->mvpDev.p, mvp_host, nm * sizeof(float), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->refDev.p, ref_u8_host, npx, cudaMemcpyHostToDevice, st));
    for (int attempt = 0;; attempt++) {
        r = ehb_render_views_fused_u8(h, mesh_ids, L, B, c->mvpDev.p, c->refDev.p, H, W, 1, nullptr, c->outDev.p, c->outDev.p + B, st);
        if (r) return r;
        CU(cudaMemcpyAsync(loss_host, c->outDev.p, B * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(g_mvp_host, c->outDev.p + B, nm * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(c->ctrHost, c->ctr, MAX_PIPES * sizeof(EhbCounters), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        { unsigned f = 0; for (int k = 0; k < MAX_PIPES; k++) f |= c->ctrHost[k].flags; if (!(f & EHB_FLAG_POOL_OVERFLOW)) break; }
        if (attempt >= 6) return fail(EHB_E_OVERFLOW, "plane pool overflow persists after growing");
        c->poolFactor *= 2.0;
        for (int k = 0; k < MAX_PIPES; k++) CU(cudaMemsetAsync(&c->ctr[k].flags, 0, sizeof(unsigned), st));
    }
    return EHB_OK;
}

int ehb_pose_compose(ehb_ctx_t h, const float* dof_dev, const float* K_dev, const float* link_poses_dev, int B, int L,
                     int H, int W, float* mvp_dev, void* stream)
{
    Ctx* c = (Ctx*)h;
    if (!c || !dof_dev || !K_dev || !link_poses_dev || !mvp_dev || B < 1 || L < 1) return fail(EHB_E_ARG, "bad pose_compose arguments");
    DeviceGuard guard(c->device);
    const int n = B * L;
    CU(launch(ehb_k_pose_compose, dim3(std::min((n + 127) / 128, 64)), dim3(128), 0, (cudaStream_t)stream, true, dof_dev, K_dev,
              link_poses_dev, n, H, W, mvp_dev));
    c->launches += 1;
    return EHB_OK;
}

int ehb_pose_backward(ehb_ctx_t h, const float* dof_dev, const float* K_dev, const float* link_poses_dev,
                      const double* g_mvp_dev, const double* loss_dev, int B, int L, int H, int W, double grad_scale,
                      double loss_scale, float* out7_dev, void* stream)
{
    Ctx* c = (Ctx*)h;
    if (!c || !dof_dev || !K_dev || !link_poses_dev || !g_mvp_dev || !loss_dev || !out7_dev || B < 1 || L < 1)
        return fail(EHB_E_ARG, "bad pose_backward arguments");
    DeviceGuard guard(c->device);
    CU(launch(ehb_k_pose_backward, dim3(1), dim3(256), 0, (cudaStream_t)stream, true, dof_dev, K_dev, link_poses_dev, g_mvp_dev,
              loss_dev, B, L, H, W, grad_scale, loss_scale, out7_dev, c->comm, 0));
    c->launches += 1;
    return EHB_OK;
}

int ehb_pose_backward_send(ehb_ctx_t h, const float* dof_dev, const float* K_dev, const float* link_poses_dev,
                           const double* g_mvp_dev, const double* loss_dev, int B, int L, int H, int W, double grad_scale,
                           double loss_scale, float* out7_dev, void* stream)
{
    Ctx* c = (Ctx*)h;
    if (!c || !dof_dev || !K_dev || !link_poses_dev || !g_mvp_dev || !loss_dev || !out7_dev || B < 1 || L < 1)
        return fail(EHB_E_ARG, "bad pose_backward arguments");
    if (!c->commReady) return fail(EHB_E_ARG, "peer mailboxes are not connected (ehb_comm_connect)");
    DeviceGuard guard(c->device);
    CU(launch(ehb_k_pose_backward, dim3(1), dim3(256), 0, (cudaStream_t)stream, true, dof_dev, K_dev, link_poses_dev, g_mvp_dev,
              loss_dev, B, L, H, W, grad_scale, loss_scale, out7_dev, c->comm, 1));
    c->launches += 1;
    return EHB_OK;
}

int ehb_adam_step(ehb_ctx_t h, float* dof_dev, const float* g7_dev, float* state_dev, float lr, float beta1, float beta2,
                  float eps, float weight_decay, float* hist_dev, int hist_cap, void* stream)
{
    Ctx* c = (Ctx*)h;
    if (!c || !dof_dev || !g7_dev || !state_dev) return fail(EHB_E_ARG, "bad adam arguments");
    DeviceGuard guard(c->device);
    CU(launch(ehb_k_adam, dim3(1), dim3(32), 0, (cudaStream_t)stream, true, dof_dev, (float*)g7_dev, state_dev, lr, beta1, beta2, eps,
              weight_decay, hist_dev, hist_cap, c->comm, 0, (const float*)nullptr, (const float*)nullptr, 0, 0, 0, (float*)nullptr));
    c->launches += 1;
    return EHB_OK;
}

int ehb_adam_step_compose(ehb_ctx_t h, float* dof_dev, float* g7_dev, float* state_dev, float lr, float beta1, float beta2,
                          float eps, float weight_decay, float* hist_dev, int hist_cap, int recv, const float* K_dev,
                          const float* link_poses_dev, int B, int L, int H, int W, float* mvp_dev, void* stream)
{
    Ctx* c = (Ctx*)h;
    if (!c || !dof_dev || !g7_dev || !state_dev || !K_dev || !link_poses_dev || !mvp_dev || B < 1 || L < 1)
        return fail(EHB_E_ARG, "bad adam_step_compose arguments");
    if (recv && !c->commReady) return fail(EHB_E_ARG, "peer mailboxes are not connected (ehb_comm_connect)");
    DeviceGuard guard(c->device);
    CU(launch(ehb_k_adam, dim3(1), dim3(256), 0, (cudaStream_t)stream, true, dof_dev, g7_dev, state_dev, lr, beta1, beta2, eps,
              weight_decay, hist_dev, hist_cap, c->comm, recv ? 1 : 0, K_dev, link_poses_dev, B * L, H, W, mvp_dev));
    c->launches += 1;
    return EHB_OK;
}

int ehb_pose_backward_adam(ehb_ctx_t h, const float* dof_dev, const float* K_dev, const float* link_poses_dev,
                           const double* g_mvp_dev, const double* loss_dev, int B, int L, int H, int W, double grad_scale,
                           double loss_scale, float* out7_dev, int exchange, float* adam_dof_dev, float* state_dev, float lr,
                           float beta1, float beta2, float eps, float weight_decay, float* hist_dev, int hist_cap,
                           float* mvp_next_dev, void* stream)
{
    Ctx* c = (Ctx*)h;
    if (!c || !dof_dev || !K_dev || !link_poses_dev || !g_mvp_dev || !loss_dev || !out7_dev || !adam_dof_dev || !state_dev ||
        B < 1 || L < 1)
        return fail(EHB_E_ARG, "bad pose_backward_adam arguments");
    if (exchange && !c->commReady) return fail(EHB_E_ARG, "peer mailboxes are not connected (ehb_comm_connect)");
    DeviceGuard guard(c->device);
    CU(launch(ehb_k_pose_adam, dim3(1), dim3(256), 0, (cudaStream_t)stream, true, dof_dev, K_dev, link_poses_dev, g_mvp_dev,
              loss_dev, B, L, H, W, grad_scale, loss_scale, out7_dev, c->comm, exchange ? 1 : 0, adam_dof_dev, state_dev, lr,
              beta1, beta2, eps, weight_decay, hist_dev, hist_cap, mvp_next_dev));
    c->launches += 1;
    return EHB_OK;
}

int ehb_adam_step_recv(ehb_ctx_t h, float* dof_dev, float* g7_dev, float* state_dev, float lr, float beta1, float beta2,
                       float eps, float weight_decay, float* hist_dev, int hist_cap, void* stream)
{
    Ctx* c = (Ctx*)h;
    if (!c || !dof_dev || !g7_dev || !state_dev) return fail(EHB_E_ARG, "bad adam arguments");
    if (!c->commReady) return fail(EHB_E_ARG, "peer mailboxes are not connected (ehb_comm_connect)");
    DeviceGuard guard(c->device);
    CU(launch(ehb_k_adam, dim3(1), dim3(32), 0, (cudaStream_t)stream, true, dof_dev, g7_dev, state_dev, lr, beta1, beta2, eps,
              weight_decay, hist_dev, hist_cap, c->comm, 1, (const float*)nullptr, (const float*)nullptr, 0, 0, 0, (float*)nullptr));
    c->launches += 1;
    return EHB_OK;
}

int ehb_solver_step_begin_u8(ehb_ctx_t h, int slot, const int* mesh_ids, int L, int B, const float* mvp_host,
                             const uint8_t* ref_u8_host, int H, int W, double* loss_host, double* g_mvp_host)
{
    Ctx* c = (Ctx*)h;
    if (!c || slot < 0 || slot >= N_SLOTS || !mvp_host || !ref_u8_host || !loss_host || !g_mvp_host) return fail(EHB_E_ARG, "bad step_begin arguments");
    if (B < 1 || L < 1) return fail(EHB_E_ARG, "empty batch");
    DeviceGuard guard(c->device);
    cudaStream_t st = c->slotStream[slot];
    int r;
    const size_t nm = (size_t)B * L * 16, no = (size_t)B + nm, npx = (size_t)B * H * W;
    if ((r = c->slotMvp[slot].ensure(nm, false))) return r;
    if ((r = c->slotOut[slot].ensure(no, false))) return r;
    if ((r = c->slotRef[slot].ensure(npx, false))) return r;
    CU(cudaMemcpyAsync(c->slotMvp[slot].p, mvp_host, nm * sizeof(float), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->slotRef[slot].p, ref_u8_host, npx, cudaMemcpyHostToDevice, st));
    Io io;
    io.ref_u8 = c->slotRef[slot].p; io.loss = c->slotOut[slot].p; io.gmvp = c->slotOut[slot].p + B;
    io.do_bwd = 1; io.clamp = 1; io.invB = 1.0f / (float)B; io.inflight = 1;
    r = run_pass(c, c->sc[MAX_PIPES + slot], mesh_ids, L, B, c->slotMvp[slot].p, H, W, EHB_MODE_FUSED, io, st);
    if (r) return r;
    CU(cudaMemcpyAsync(loss_host, c->slotOut[slot].p, B * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(g_mvp_host, c->slotOut[slot].p + B, nm * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(c->ctrHost + MAX_PIPES + slot, c->ctr + MAX_PIPES + slot, sizeof(EhbCounters), cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(c->slotDone[slot], st));
    return EHB_OK;
}


int ehb_solver_step_begin_ref(ehb_ctx_t h, int slot, const int* mesh_ids, int L, int B, const float* mvp_host, int ref_id,
                              int first_view, int H, int W, double* loss_host, double* g_mvp_host)
{
    Ctx* c = (Ctx*)h;
    if (!c || slot < 0 || slot >= N_SLOTS || !mvp_host || !loss_host || !g_mvp_host) return fail(EHB_E_ARG, "bad step_begin arguments");
    if (B < 1 || L < 1) return fail(EHB_E_ARG, "empty batch");
    DeviceGuard guard(c->device);
    cudaStream_t st = c->slotStream[slot];
    int r;
    const size_t nm = (size_t)B * L * 16, no = (size_t)B + nm;
    if ((r = c->slotMvp[slot].ensure(nm, false))) return r;
    if ((r = c->slotOut[slot].ensure(no, false))) return r;
    Io io;
    if ((r = ref_io(c, ref_id, first_view, B, H, W, io))) return r;
    CU(cudaMemcpyAsync(c->slotMvp[slot].p, mvp_host, nm * sizeof(float), cudaMemcpyHostToDevice, st));
    io.loss = c->slotOut[slot].p; io.gmvp = c->slotOut[slot].p + B;
    io.do_bwd = 1; io.clamp = 1; io.invB = 1.0f / (float)B; io.inflight = 1;
    r = run_pass(c, c->sc[MAX_PIPES + slot], mesh_ids, L, B, c->slotMvp[slot].p, H, W, EHB_MODE_FUSED, io, st);
    if (r) return r;
    CU(cudaMemcpyAsync(loss_host, c->slotOut[slot].p, B * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(g_mvp_host, c->slotOut[slot].p + B, nm * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(c->ctrHost + MAX_PIPES + slot, c->ctr + MAX_PIPES + slot, sizeof(EhbCounters), cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(c->slotDone[slot], st));
    return EHB_OK;
}

int ehb_step_begin(ehb_ctx_t h, int slot, const int* mesh_ids, int L, int B, int ref_id, int first_view, int H, int W,
                   const ehb_step_io_t* sio)
{
    Ctx* c = (Ctx*)h;
    if (!c || slot < 0 || slot >= N_SLOTS || !mesh_ids || !sio || (!sio->mvp_host && !sio->mvp_dev)) return fail(EHB_E_ARG, "bad step_begin arguments");
    if (B < 1 || L < 1) return fail(EHB_E_ARG, "empty batch");
    const bool adam = sio->adam_dof_dev != nullptr;
    const bool pose = sio->out7_dev || sio->out7_host || adam;
    if (pose && (!sio->dof_dev || !sio->K_dev || !sio->link_poses_dev)) return fail(EHB_E_ARG, "the pose chain needs dof_dev, K_dev and link_poses_dev");
    if (adam && !sio->adam_state_dev) return fail(EHB_E_ARG, "the Adam update needs adam_state_dev");
    const int exch = sio->exchange ? 1 : 0;
    if (exch && !c->commReady) return fail(EHB_E_ARG, "peer mailboxes are not connected (ehb_comm_connect)");
    if (exch && !adam) return fail(EHB_E_ARG, "an exchanged step ends with the Adam update (it receives the sum)");
    DeviceGuard guard(c->device);
    cudaStream_t st = c->slotStream[slot];
    int r;
    const size_t nm = (size_t)B * L * 16, no = (size_t)B + nm + 8;
    if ((r = c->slotOut[slot].ensure(no, false))) return r;
    Io io;
    if ((r = ref_io(c, ref_id, first_view, B, H, W, io))) return r;
    const float* mvp = sio->mvp_dev;
    if (!mvp) {
        if ((r = c->slotMvp[slot].ensure(nm, false))) return r;
        CU(cudaMemcpyAsync(c->slotMvp[slot].p, sio->mvp_host, nm * sizeof(float), cudaMemcpyHostToDevice, st));
        mvp = c->slotMvp[slot].p;
    }
    io.masks = sio->masks_dev; io.loss = c->slotOut[slot].p; io.gmvp = c->slotOut[slot].p + B;
    io.do_bwd = 1; io.clamp = 1; io.invB = 1.0f / (float)B; io.inflight = 1;
    r = run_pass(c, c->sc[MAX_PIPES + slot], mesh_ids, L, B, mvp, H, W, EHB_MODE_FUSED, io, st);
    if (r) return r;
    if (pose) {
        float* o7 = sio->out7_dev ? sio->out7_dev : reinterpret_cast<float*>(c->slotOut[slot].p + B + nm);
        const double gs = sio->grad_scale != 0.0 ? sio->grad_scale : 1.0, ls = sio->loss_scale != 0.0 ? sio->loss_scale : 1.0 / (double)B;
        if (adam) {   // pose chain, exchange and Adam in one launch
            CU(launch(ehb_k_pose_adam, dim3(1), dim3(256), 0, st, true, sio->dof_dev, sio->K_dev, sio->link_poses_dev,
                      (const double*)(c->slotOut[slot].p + B), (const double*)c->slotOut[slot].p, B, L, H, W, gs, ls, o7,
                      c->commSlot[slot], exch, sio->adam_dof_dev, sio->adam_state_dev, sio->lr, 0.9f, 0.999f, 1e-8f,
                      sio->weight_decay, (float*)nullptr, 0, (float*)nullptr));
        } else {
            CU(launch(ehb_k_pose_backward, dim3(1), dim3(256), 0, st, true, sio->dof_dev, sio->K_dev, sio->link_poses_dev,
                      (const double*)(c->slotOut[slot].p + B), (const double*)c->slotOut[slot].p, B, L, H, W, gs, ls, o7,
                      c->commSlot[slot], exch));
        }
        c->launches += 1;
        if (sio->out7_host) CU(cudaMemcpyAsync(sio->out7_host, o7, 7 * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    if (sio->loss_host) CU(cudaMemcpyAsync(sio->loss_host, c->slotOut[slot].p, B * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (sio->g_mvp_host) CU(cudaMemcpyAsync(sio->g_mvp_host, c->slotOut[slot].p + B, nm * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(c->ctrHost + MAX_PIPES + slot, c->ctr + MAX_PIPES + slot, sizeof(EhbCounters), cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(c->slotDone[slot], st));
    return EHB_OK;
}

int ehb_slot_stream(ehb_ctx_t h, int slot, void** stream)
{
    Ctx* c = (Ctx*)h;
    if (!c || slot < 0 || slot >= N_SLOTS || !stream) return fail(EHB_E_ARG, "bad slot");
    *stream = (void*)c->slotStream[slot];
    return EHB_OK;
}

int ehb_slots_fork(ehb_ctx_t h, void* stream)
{
    Ctx* c = (Ctx*)h;
    if (!c) return fail(EHB_E_ARG, "null context");
    DeviceGuard guard(c->device);
    CU(cudaEventRecord(c->evFork, (cudaStream_t)stream));
    for (int k = 0; k < N_SLOTS; k++) CU(cudaStreamWaitEvent(c->slotStream[k], c->evFork, 0));
    return EHB_OK;
}

int ehb_slots_join(ehb_ctx_t h, void* stream)
{
    Ctx* c = (Ctx*)h;
    if (!c) return fail(EHB_E_ARG, "null context");
    DeviceGuard guard(c->device);
    for (int k = 0; k < N_SLOTS; k++) {
        CU(cudaEventRecord(c->slotDone[k], c->slotStream[k]));
        CU(cudaStreamWaitEvent((cudaStream_t)stream, c->slotDone[k], 0));
    }
    return EHB_OK;
}

int ehb_solver_step_end(ehb_ctx_t h, int slot)
{
    Ctx* c = (Ctx*)h;
    if (!c || slot < 0 || slot >= N_SLOTS) return fail(EHB_E_ARG, "bad slot");
    DeviceGuard guard(c->device);
    CU(cudaEventSynchronize(c->slotDone[slot]));
    if (c->ctrHost[MAX_PIPES + slot].flags & EHB_FLAG_POOL_OVERFLOW) {
        c->poolFactor *= 2.0;
        CU(cudaMemsetAsync(&c->ctr[MAX_PIPES + slot].flags, 0, sizeof(unsigned), c->slotStream[slot]));
        return fail(EHB_E_OVERFLOW, "depth-plane pool was too small for this step; it has been grown, submit the step again");
    }
    return EHB_OK;
}

int ehb_comm_local_handle(ehb_ctx_t h, void* handle64)
{
    Ctx* c = (Ctx*)h;
    if (!c || !handle64) return fail(EHB_E_ARG, "null pointer argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DeviceGuard guard(c->device);
    if (!c->commBox) {
        const size_t words = (size_t)(1 + N_SLOTS) * COMM_CH_WORDS;
        CU(cudaMalloc((void**)&c->commBox, words * sizeof(unsigned)));
        CU(cudaMemset(c->commBox, 0, words * sizeof(unsigned)));
        CU(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t hd;
    CU(cudaIpcGetMemHandle(&hd, c->commBox));
    memcpy(handle64, &hd, 64);
    return EHB_OK;
}

int ehb_comm_connect(ehb_ctx_t h, int rank, int world, const void* handles)
{
    Ctx* c = (Ctx*)h;
    if (!c || !handles || world < 1 || world > EHB_COMM_MAX || rank < 0 || rank >= world || !c->commBox)
        return fail(EHB_E_ARG, "bad comm arguments (call ehb_comm_local_handle first; world <= %d)", EHB_COMM_MAX);
    DeviceGuard guard(c->device);
    for (int r = 0; r < world; r++) {
        if (r == rank) { c->comm.peer[r] = c->commBox; continue; }
        cudaIpcMemHandle_t hd;
        memcpy(&hd, (const char*)handles + 64 * r, 64);
        void* ptr = nullptr;
        CU(cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
        c->comm.peer[r] = (unsigned int*)ptr;
    }
    c->comm.step = c->commBox + 2 * EHB_COMM_MAX * 8;
    c->comm.rank = rank; c->comm.world = world;
    for (int k = 0; k < N_SLOTS; k++) {   // channel 1 + k: the same layout, COMM_CH_WORDS further on in every mailbox
        c->commSlot[k] = c->comm;
        for (int r = 0; r < world; r++) c->commSlot[k].peer[r] = c->comm.peer[r] + (size_t)(1 + k) * COMM_CH_WORDS;
        c->commSlot[k].step = c->comm.step + (size_t)(1 + k) * COMM_CH_WORDS;
    }
    c->commReady = true;
    return EHB_OK;
}

int ehb_allreduce7(ehb_ctx_t h, float* g7_dev, void* stream)
{
    Ctx* c = (Ctx*)h;
    if (!c || !g7_dev) return fail(EHB_E_ARG, "null pointer argument");
    if (!c->commReady) return fail(EHB_E_ARG, "peer mailboxes are not connected (ehb_comm_connect)");
    DeviceGuard guard(c->device);
    ehb_k_allreduce7<<<1, 32, 0, (cudaStream_t)stream>>>(c->comm, g7_dev);
    c->launches += 1;
    CU(cudaGetLastError());
    return EHB_OK;
}

long long ehb_launch_count(ehb_ctx_t h) { return h ? ((Ctx*)h)->launches : 0; }

}  // extern "C"
