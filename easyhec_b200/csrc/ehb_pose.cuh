// ehb_pose.cuh -- the 6-DoF pose chain around the rasterizer, as three tiny kernels so that one optimizer
// iteration is a fixed launch sequence (CUDA-graph capturable, no host round trip):
//   k_pose_compose : dof -> Tc_c2b = se3_exp_map(dof)^T (easyhec/utils/pytorch3d_se3.py:46-130, transposed as in
//                    rb_solver.py:52) -> mvp[b,l] = K_to_projection(K) @ diag(1,-1,-1,1) @ Tc_c2b @ link_poses[b,l]
//                    (easyhec/utils/nvdiffrast_utils.py:5-11, nvdiffrast_renderer.py:33-37, rb_solver.py:63)
//   k_pose_backward: d loss / d mvp[b,l] -> d loss / d Tc_c2b -> d loss / d dof (forward-mode dual numbers through the
//                    same exp map, which reproduces autograd including the eps clamp), plus the mean loss
//   k_adam         : torch.optim.Adam with L2 weight decay on dof (easyhec/solver/build.py:12-29), step counter and
//                    the dof history (rb_solver.py:50-51) kept on the device.
#pragma once
#include <cuda_runtime.h>

// programmatic dependent launch: see ehb_pdl_enter (ehb_kernels.cuh); the pose kernels are chained the same way
__device__ __forceinline__ void ehb_pose_pdl_enter()
{
#ifdef EHB_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

struct EhbDual { double v, d; };
__device__ __forceinline__ EhbDual operator+(EhbDual a, EhbDual b) { return {a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ EhbDual operator-(EhbDual a, EhbDual b) { return {a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ EhbDual operator*(EhbDual a, EhbDual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
__device__ __forceinline__ EhbDual operator/(EhbDual a, EhbDual b) { return {a.v / b.v, (a.d * b.v - a.v * b.d) / (b.v * b.v)}; }
__device__ __forceinline__ EhbDual ehb_sin(EhbDual a) { return {sin(a.v), cos(a.v) * a.d}; }
__device__ __forceinline__ EhbDual ehb_cos(EhbDual a) { return {cos(a.v), -sin(a.v) * a.d}; }
__device__ __forceinline__ EhbDual ehb_sqrt(EhbDual a) { const double s = sqrt(a.v); return {s, a.d / (2.0 * s)}; }
__device__ __forceinline__ EhbDual ehb_clamp_min(EhbDual a, double lo) { return a.v < lo ? EhbDual{lo, 0.0} : a; }
__device__ __forceinline__ EhbDual ehb_const(EhbDual, double c) { return {c, 0.0}; }
__device__ __forceinline__ float ehb_sin(float a) { return sinf(a); }
__device__ __forceinline__ float ehb_cos(float a) { return cosf(a); }
__device__ __forceinline__ float ehb_sqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ float ehb_clamp_min(float a, double lo) { return fmaxf(a, (float)lo); }
__device__ __forceinline__ float ehb_const(float, double c) { return (float)c; }

// dof = [t | w] -> top three rows of Tc_c2b (3x4, row-major): [R | V t]
template <typename T>
__device__ void ehb_se3_exp(const T* dof, T* out, double eps)
{
    const T one = ehb_const(dof[0], 1.0), zero = ehb_const(dof[0], 0.0);
    const T wx = dof[3], wy = dof[4], wz = dof[5];
    const T th = ehb_sqrt(ehb_clamp_min(wx * wx + wy * wy + wz * wz, eps));
    const T inv = one / th;
    const T s = ehb_sin(th), c = ehb_cos(th);
    const T f1 = inv * s;
    const T f2 = inv * inv * (one - c);
    const T f3 = (th - s) / (th * th * th);
    const T Kx[9] = {zero, zero - wz, wy, wz, zero, zero - wx, zero - wy, wx, zero};
    T K2[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) K2[3 * i + j] = Kx[3 * i] * Kx[j] + Kx[3 * i + 1] * Kx[3 + j] + Kx[3 * i + 2] * Kx[6 + j];
    for (int i = 0; i < 3; i++) {
        T tr = zero;
        for (int j = 0; j < 3; j++) {
            const T id = (i == j) ? one : zero;
            out[4 * i + j] = f1 * Kx[3 * i + j] + f2 * K2[3 * i + j] + id;
            const T Vij = id + Kx[3 * i + j] * f2 + K2[3 * i + j] * f3;
            tr = tr + Vij * dof[j];
        }
        out[4 * i + 3] = tr;
    }
}

// P = K_to_projection(K, H, W, n = 0.001, f = 10) @ diag(1,-1,-1,1), row-major 4x4
__device__ __forceinline__ void ehb_projection(const float* K, int H, int W, float* P)
{
    const float fu = K[0], fv = K[4], cu = K[2], cv = K[5];
    const float a = (float)(-(10.0 + 0.001) / (10.0 - 0.001)), b = (float)(-2.0 * 10.0 * 0.001 / (10.0 - 0.001));
    for (int i = 0; i < 16; i++) P[i] = 0.f;
    P[0] = (2.f * fu) / (float)W;
    P[2] = -((-2.f * cu) / (float)W + 1.f);
    P[5] = -((2.f * fv) / (float)H);
    P[6] = -((2.f * cv) / (float)H - 1.f);
    P[10] = -a;
    P[11] = b;
    P[14] = 1.f;
}

// mvp[i] = P @ (Tc @ lp[i]) for the matrices i = first, first + stride, ... (all threads of a block; M, Tc: 16 floats each in
// shared memory, thread 0 fills them from dof / K first)
__device__ __forceinline__ void ehb_compose_block(const float* __restrict__ dof, const float* __restrict__ K,
                                                  const float* __restrict__ lp, int n, int H, int W, float* __restrict__ mvp,
                                                  float* M, float* Tc, int first, int stride)
{
    if (threadIdx.x == 0) {
        float d[6], t[12];
        for (int i = 0; i < 6; i++) d[i] = dof[i];
        ehb_se3_exp<float>(d, t, 1e-4);
        for (int i = 0; i < 12; i++) Tc[i] = t[i];
        Tc[12] = 0.f; Tc[13] = 0.f; Tc[14] = 0.f; Tc[15] = 1.f;
        float k[9];
        for (int i = 0; i < 9; i++) k[i] = K[i];
        ehb_projection(k, H, W, M);
    }
    __syncthreads();
    for (int i = first; i < n; i += stride) {
        float A[16], B[16];
        for (int j = 0; j < 16; j++) A[j] = lp[(size_t)i * 16 + j];
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) {
                float s = Tc[4 * r] * A[c];
                s = __fmaf_rn(Tc[4 * r + 1], A[4 + c], s);
                s = __fmaf_rn(Tc[4 * r + 2], A[8 + c], s);
                s = __fmaf_rn(Tc[4 * r + 3], A[12 + c], s);
                B[4 * r + c] = s;
            }
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) {
                float s = M[4 * r] * B[c];
                s = __fmaf_rn(M[4 * r + 1], B[4 + c], s);
                s = __fmaf_rn(M[4 * r + 2], B[8 + c], s);
                s = __fmaf_rn(M[4 * r + 3], B[12 + c], s);
                mvp[(size_t)i * 16 + 4 * r + c] = s;
            }
    }
}

__global__ void ehb_k_pose_compose(const float* __restrict__ dof, const float* __restrict__ K,
                                   const float* __restrict__ lp, int n, int H, int W, float* __restrict__ mvp)
{
    ehb_pose_pdl_enter();
    __shared__ float M[16];   // P @ [Tc; 0 0 0 1] is NOT pre-multiplied: the reference composes P @ (Tc @ lp)
    __shared__ float Tc[16];
    ehb_compose_block(dof, K, lp, n, H, W, mvp, M, Tc, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// ---------------------------------------------------------------------------------------------------------------
// One-shot all-reduce of the 7 floats (d loss/d dof, loss) over NVLink peer memory.  Every rank owns a mailbox
// [2 parities][world][8 words] that its peers map with CUDA IPC.  One step: write (7 floats, step tag) into slot
// [parity][my rank] of EVERY rank's mailbox (plain stores over NVLink / NVSwitch), fence, then spin until all slots of
// the own mailbox carry this step's tag and add them in rank order (so every rank gets bit-identical sums).
// Two parities are enough: a rank can only be one step ahead of a peer that has not yet read its previous message.
// 28 bytes cross the switch per peer, so the cost is one NVLink round trip (~2-4 us) instead of a collective launch.
#define EHB_COMM_MAX 16
struct EhbComm {
    unsigned int* peer[EHB_COMM_MAX];   // peer[r] = rank r's mailbox, as mapped in this process (peer[rank] = own)
    unsigned int* step;                 // device counter of completed all-reduces (own)
    int rank, world;
};

// one block of 256 threads.  out7 = { d loss / d dof [6], loss } with the caller's scales applied.  The tail is spread over
// threads (12 for P^T S, 6 for the six dual-number evaluations of the exp map): fp64 on one thread was 6 us of pure latency.
struct EhbPoseShared {
    float out[8];
    double S[8][17];
    double T[17];
    double G[12];
    double dT[6][12];        // d exp_map(dof)[j] / d dof[i]: depends on dof only, computed beside the reduction
};
__device__ __forceinline__ void ehb_pose_backward_block(EhbPoseShared& sh, const float* __restrict__ dof, const float* __restrict__ K,
                                                        const float* __restrict__ lp, const double* __restrict__ gmvp,
                                                        const double* __restrict__ loss, int B, int L, int H, int W,
                                                        double grad_scale, double loss_scale, float* __restrict__ out7,
                                                        const EhbComm& cm, int send)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // the six dual-number evaluations of the exp map need dof only: the last warp does them while the others wait for the
    // gradient's loads (they were 1 - 2 us at the end of the chain)
    if (warp == 7 && lane < 6) {
        EhbDual d[6], t[12];
        for (int j = 0; j < 6; j++) d[j] = {(double)dof[j], j == lane ? 1.0 : 0.0};
        ehb_se3_exp<EhbDual>(d, t, 1e-4);
        for (int j = 0; j < 12; j++) sh.dT[lane][j] = t[j].d;
    }
    // S = sum_{b,l} g_mvp[b,l] @ lp[b,l]^T ;  S[16] = sum_b loss_b
    double acc[17];
    for (int i = 0; i < 17; i++) acc[i] = 0.0;
    for (int i = tid; i < B * L; i += 256) {
        const double* g = gmvp + (size_t)i * 16;
        const float* a = lp + (size_t)i * 16;
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++) {
                double s = 0.0;
                for (int k = 0; k < 4; k++) s += g[4 * r + k] * (double)a[4 * c + k];
                acc[4 * r + c] += s;
            }
    }
    for (int i = tid; i < B; i += 256) acc[16] += loss[i];
    for (int i = 0; i < 17; i++) {
        double v = acc[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sh.S[warp][i] = v;
    }
    __syncthreads();
    if (tid < 17) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += sh.S[w][tid];
        sh.T[tid] = t;
    }
    __syncthreads();
    if (tid < 12) {   // top three rows of P^T @ S
        float k[9], P[16];
        for (int i = 0; i < 9; i++) k[i] = K[i];
        ehb_projection(k, H, W, P);
        const int r = tid >> 2, c = tid & 3;
        double s = 0.0;
        for (int q = 0; q < 4; q++) s += (double)P[4 * q + r] * sh.T[4 * q + c];
        sh.G[tid] = s;
    }
    __syncthreads();
    if (tid < 6) {
        double s = 0.0;
        for (int j = 0; j < 12; j++) s += sh.G[j] * sh.dT[tid][j];
        out7[tid] = sh.out[tid] = (float)(s * grad_scale);
    } else if (tid == 6) {
        out7[6] = sh.out[6] = (float)(sh.T[16] * loss_scale);
    }
    if (send) {
        // first half of the all-reduce, fused: this rank's 7 floats + the step tag go into every peer's mailbox (plain
        // stores over NVLink); the Adam half (recv) waits for all of them and adds them in rank order
        __syncthreads();
        if (tid < cm.world) {
            const unsigned int step = *cm.step + 1u;
            volatile unsigned int* dst = cm.peer[tid] + ((size_t)(step & 1u) * EHB_COMM_MAX + cm.rank) * 8;
            for (int i = 0; i < 7; i++) dst[i] = __float_as_uint(sh.out[i]);
            __threadfence_system();
            dst[7] = step;
        }
    }
}

__global__ void __launch_bounds__(256) ehb_k_pose_backward(const float* __restrict__ dof, const float* __restrict__ K,
                                                           const float* __restrict__ lp, const double* __restrict__ gmvp,
                                                           const double* __restrict__ loss, int B, int L, int H, int W,
                                                           double grad_scale, double loss_scale, float* __restrict__ out7,
                                                           const EhbComm cm, int send)
{
    ehb_pose_pdl_enter();
    __shared__ EhbPoseShared sh;
    ehb_pose_backward_block(sh, dof, K, lp, gmvp, loss, B, L, H, W, grad_scale, loss_scale, out7, cm, send);
}

// state = { m[6], v[6], t }.  torch.optim.Adam (L2 weight decay folded into the gradient, bias-corrected).
// With mvp_out != NULL the same launch composes the matrices of the NEXT iteration from the updated parameters
// (ehb_k_pose_compose's arithmetic): one launch less on the critical path of a pose-optimisation iteration.
struct EhbAdamShared {
    float M[16], Tc[16];
    float val[EHB_COMM_MAX][7];
};
__device__ __forceinline__ void ehb_adam_block(EhbAdamShared& sh, float* __restrict__ dof, float* __restrict__ g7,
                                               float* __restrict__ state, float lr, float beta1, float beta2, float eps, float wd,
                                               float* __restrict__ hist, int hist_cap, const EhbComm& cm, int recv,
                                               const float* __restrict__ K, const float* __restrict__ lp, int n, int H, int W,
                                               float* __restrict__ mvp_out, const double* bc = nullptr)
{
    if (recv) {
        // second half of the fused all-reduce: wait until every rank's message of this step is in the own mailbox, add
        // them in rank order (bit-identical sums on every rank), leave the sum in g7
        const int t = threadIdx.x;
        const unsigned int step = *cm.step + 1u;
        if (t < cm.world) {
            volatile unsigned int* src = cm.peer[cm.rank] + ((size_t)(step & 1u) * EHB_COMM_MAX + t) * 8;
            while (src[7] != step) { }
            __threadfence_system();
            for (int i = 0; i < 7; i++) sh.val[t][i] = __uint_as_float(src[i]);
        }
        __syncthreads();
        if (t < 7) {
            float acc = 0.f;
            for (int r = 0; r < cm.world; r++) acc += sh.val[r][t];
            g7[t] = acc;
        }
        if (t == 0) *cm.step = step;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int t = (int)state[12] + 1;
        if (hist && t - 1 < hist_cap)
            for (int i = 0; i < 6; i++) hist[(size_t)(t - 1) * 6 + i] = dof[i];
        // (bc: the bias corrections of step t, computed by idle threads at the start of the launch)
        const double bc1 = bc ? bc[0] : 1.0 - pow((double)beta1, (double)t), bc2 = bc ? bc[1] : 1.0 - pow((double)beta2, (double)t);
        for (int i = 0; i < 6; i++) {
            float g = g7[i];
            if (wd != 0.f) g = g + wd * dof[i];
            const float m = beta1 * state[i] + (1.f - beta1) * g;
            const float v = beta2 * state[6 + i] + (1.f - beta2) * g * g;
            state[i] = m; state[6 + i] = v;
            const float step_size = (float)((double)lr / bc1);
            const float denom = (float)((double)sqrtf(v) / sqrt(bc2)) + eps;
            dof[i] = dof[i] - step_size * (m / denom);
        }
        state[12] = (float)t;
    }
    if (mvp_out) {
        __syncthreads();                                   // (thread 0's parameter update is visible to itself; the others wait)
        ehb_compose_block(dof, K, lp, n, H, W, mvp_out, sh.M, sh.Tc, threadIdx.x, blockDim.x);
    }
}

__global__ void ehb_k_adam(float* __restrict__ dof, float* __restrict__ g7, float* __restrict__ state, float lr,
                           float beta1, float beta2, float eps, float wd, float* __restrict__ hist, int hist_cap,
                           const EhbComm cm, int recv, const float* __restrict__ K, const float* __restrict__ lp, int n,
                           int H, int W, float* __restrict__ mvp_out)
{
    ehb_pose_pdl_enter();
    if (blockIdx.x != 0) return;
    __shared__ EhbAdamShared sh;
    ehb_adam_block(sh, dof, g7, state, lr, beta1, beta2, eps, wd, hist, hist_cap, cm, recv, K, lp, n, H, W, mvp_out);
}

// The tail of a solver iteration in ONE launch: pose chain (d loss/d mvp -> the 7 floats, + the mailbox send), then -- in the
// same block, behind a barrier -- the exchange's receive, Adam on `adam_dof` and optionally the next iteration's matrices.
// A launch less on the critical path of every iteration (3 us of a 75 us iteration at 640x480).
__global__ void __launch_bounds__(256) ehb_k_pose_adam(const float* dof /* may alias adam_dof */, const float* __restrict__ K,
                                                       const float* __restrict__ lp, const double* __restrict__ gmvp,
                                                       const double* __restrict__ loss, int B, int L, int H, int W,
                                                       double grad_scale, double loss_scale, float* __restrict__ out7,
                                                       const EhbComm cm, int exch, float* adam_dof,
                                                       float* __restrict__ state, float lr, float beta1, float beta2, float eps,
                                                       float wd, float* __restrict__ hist, int hist_cap,
                                                       float* __restrict__ mvp_out)
{
    ehb_pose_pdl_enter();
    __shared__ EhbPoseShared shp;
    __shared__ EhbAdamShared sha;
    __shared__ double s_bc[2];
    if (threadIdx.x == 230 || threadIdx.x == 231) {        // Adam's bias corrections need the step count only
        const int k = threadIdx.x - 230;
        s_bc[k] = 1.0 - pow((double)(k ? beta2 : beta1), (double)((int)state[12] + 1));
    }
    ehb_pose_backward_block(shp, dof, K, lp, gmvp, loss, B, L, H, W, grad_scale, loss_scale, out7, cm, exch);
    __syncthreads();                                       // out7 as written by this block; dof read before Adam changes it
    ehb_adam_block(sha, adam_dof, out7, state, lr, beta1, beta2, eps, wd, hist, hist_cap, cm, exch, K, lp, B * L, H, W, mvp_out, s_bc);
}

__global__ void ehb_k_allreduce7(const EhbComm cm, float* __restrict__ g7)
{
    const int t = threadIdx.x;
    const unsigned int step = *cm.step + 1u;
    const unsigned int parity = step & 1u;
    if (t < cm.world) {
        volatile unsigned int* dst = cm.peer[t] + ((size_t)parity * EHB_COMM_MAX + cm.rank) * 8;
        for (int i = 0; i < 7; i++) dst[i] = __float_as_uint(g7[i]);
        __threadfence_system();
        dst[7] = step;
    }
    __syncthreads();
    __shared__ float s_val[EHB_COMM_MAX][7];
    if (t < cm.world) {
        volatile unsigned int* src = cm.peer[cm.rank] + ((size_t)parity * EHB_COMM_MAX + t) * 8;
        while (src[7] != step) { }
        __threadfence_system();
        for (int i = 0; i < 7; i++) s_val[t][i] = __uint_as_float(src[i]);
    }
    __syncthreads();
    if (t < 7) {
        float acc = 0.f;
        for (int r = 0; r < cm.world; r++) acc += s_val[r][t];
        g7[t] = acc;
    }
    if (t == 0) *cm.step = step;
}
