/*
 * easyhec_b200.h -- C ABI of the B200-native silhouette rasterizer (libehb.so).
 *
 * Drop-in boundary for EasyHeC's render_mask hot path.  Every entry point takes plain pointers,
 * sizes and a CUDA stream handle (void* = cudaStream_t); no torch types.  All functions return 0
 * on success or a negative EHB_E_* code; ehb_last_error() gives the message of the last failure
 * on the calling thread.  Nothing here falls back to the CPU: without a usable CUDA device every
 * compute entry point fails with EHB_E_CUDA.
 *
 * Reference interfaces replaced (paths relative to the EasyHeC tree):
 *   dr.RasterizeCudaContext()                 easyhec/structures/nvdiffrast_renderer.py:23   -> ehb_ctx_create
 *   dr.rasterize / interpolate / antialias    easyhec/structures/nvdiffrast_renderer.py:39-47 -> ehb_render_mask_fwd / _bwd
 *   batch_render_mask (packed links, no AA)   easyhec/structures/nvdiffrast_renderer.py:50-73,
 *                                             easyhec/utils/render_api.py:70-96              -> ehb_render_binary_batch
 *   RBSolver per-view / per-link loop + loss  easyhec/modeling/models/rb_solve/rb_solver.py:60-72 -> ehb_render_views_fused
 *   SpaceExplorer variance score              easyhec/modeling/models/rb_solve/space_explorer.py:152-165 -> ehb_variance_score,
 *                                                                                                ehb_explore_scores
 *
 * Conventions:
 *   mvp        row-major 4x4 fp32, clip = mvp * [x y z 1]^T, mvp = K_to_projection(K,H,W) @ diag(1,-1,-1,1) @ pose
 *              (easyhec/utils/nvdiffrast_utils.py:5-18, nvdiffrast_renderer.py:33-37)
 *   masks      image rows, row 0 = top (the reference flips nvdiffrast's output, nvdiffrast_renderer.py:47)
 *   *_dev      device pointer on the context's device;   *_host   host pointer
 *   stream     work is only enqueued; nothing synchronises except where stated.  One context per
 *              device, not thread-safe, one stream at a time (same contract as the reference ctx).
 */
#ifndef EASYHEC_B200_H
#define EASYHEC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define EHB_API __attribute__((visibility("default")))
#else
#define EHB_API
#endif

#define EHB_OK 0
#define EHB_E_ARG (-1)      /* bad argument (null pointer, size, unknown mesh id, resolution > 8160) */
#define EHB_E_CUDA (-2)     /* CUDA runtime error or no device */
#define EHB_E_CAPACITY (-3) /* scratch too small and cannot grow here (stream capture in progress) */
#define EHB_E_OVERFLOW (-4) /* the depth-plane pool stayed too small after growing (see ehb_ctx_status) */

/* bits of *flags from ehb_ctx_status */
#define EHB_FLAG_POOL_OVERFLOW 1u /* a scratch pool (depth planes, jobs, silhouette pairs) was too small: results of that
                                     launch are incomplete */
#define EHB_FLAG_PAIR_OVERFLOW EHB_FLAG_POOL_OVERFLOW
#define EHB_FLAG_NEEDS_CLIP 2u    /* informational: triangles left the depth range / guard band and went through the
                                     frustum clipper (drawn; ehb_ctx_status reports how many) */
#define EHB_FLAG_QUEUES_FULL 4u   /* informational: the deferred-triangle queues were full, some large triangles were drawn
                                     inline (slower, results complete) */

typedef void* ehb_ctx_t;

EHB_API int ehb_version(void);
EHB_API const char* ehb_last_error(void);

/* Context = per-device scratch (depth-plane pool, tile queue) + registered meshes.  Replaces dr.RasterizeCudaContext. */
EHB_API int ehb_ctx_create(int device, ehb_ctx_t* out);
EHB_API int ehb_ctx_destroy(ehb_ctx_t ctx);
/* Pre-size scratch so later launches never allocate (required before CUDA-graph capture).
 * n_items = views (or renders) per launch, max_faces = sum of faces over the links of one item. */
EHB_API int ehb_ctx_reserve(ehb_ctx_t ctx, int n_items, int n_links, int max_faces, int H, int W);
/* Tie rule for a pixel centre lying exactly on a snapped edge: 0 (default) or 1 (mirror); see DESIGN.md. */
EHB_API int ehb_ctx_set_fill_rule(ehb_ctx_t ctx, int rule);
/* A call's items (views / renders) are split over n (1..4) independent pipelines that run concurrently on internal streams
 * forked from, and joined back into, the caller's stream (CUDA-graph capturable).  Default: 3 for calls of more than 16
 * items, 1 otherwise (measured); setting n makes every call of >= 2 items use min(n, items) pipelines. */
EHB_API int ehb_ctx_set_pipelines(ehb_ctx_t ctx, int n);
/* Bytes of depth-plane pool a pipeline may reserve up front (default 8e9).  While the worst case of a call
 * (items x links x H x W x 8 B) fits, the pool cannot overflow; beyond it the pool starts at 2 screens per item and
 * EHB_FLAG_POOL_OVERFLOW asks for ehb_ctx_grow_scratch + a rerun. */
EHB_API int ehb_ctx_set_pool_budget(ehb_ctx_t ctx, double bytes);
/* Doubles the scratch pools used for later launches (call after EHB_FLAG_POOL_OVERFLOW). */
EHB_API int ehb_ctx_grow_scratch(ehb_ctx_t ctx);
/* Per-kernel timing for benchmarks: while enabled every pass records CUDA events around its four kernels on the
 * caller's stream (and runs it as a single pipeline).  ehb_ctx_kernel_times synchronises, returns the summed
 * milliseconds of the four stages {table, front, raster (+ raster_big), tiles} (ms4[4]) over the passes recorded since
 * the last query, and their number. */
EHB_API int ehb_ctx_profile(ehb_ctx_t ctx, int enable);
EHB_API int ehb_ctx_kernel_times(ehb_ctx_t ctx, double* ms4, long long* n_passes);
/* the same without forgetting the recorded passes (for passes captured in a CUDA graph: every replay records their events
 * again, so the times of the last replay can be read after each one) */
EHB_API int ehb_ctx_kernel_times_peek(ehb_ctx_t ctx, double* ms4, long long* n_passes);
/* Per-kernel timing covers the stages {table, front, raster (+ raster_big), tiles (= windows + compose + pairgrad)}. */
/* Developer aid: 16 raw 64-bit scratch counters of the first pipeline (zero unless a debug build fills them). */
EHB_API int ehb_ctx_debug_counters(ehb_ctx_t ctx, unsigned long long* out16, int reset);
/* Developer aid: first call allocates a device scratch buffer that debug builds may fill, later calls copy up to
 * n_words 64-bit words of it out. */
EHB_API int ehb_ctx_debug_buffer(ehb_ctx_t ctx, unsigned long long* out, int n_words);
/* developer builds (-DEHB_MARKS): progress marks the kernels leave in host-visible memory, readable while a kernel runs */
EHB_API int ehb_ctx_debug_marks(ehb_ctx_t ctx, unsigned* out16);
/* Synchronises the device, returns and clears the sticky flags, reports the triangles that went through the clipper. */
EHB_API int ehb_ctx_status(ehb_ctx_t ctx, unsigned* flags, long long* n_need_clip);
/* The same sticky flags WITHOUT synchronising: the kernels also set them in host-visible (mapped) memory, so a caller can
 * look between launches or graph replays at no cost.  A flag raised by work that is still running shows up on a later
 * poll; ehb_ctx_status clears what both report. */
EHB_API int ehb_ctx_poll(ehb_ctx_t ctx, unsigned* flags);

/* Register a mesh once: verts_host f32[V*3], faces_host i32[F*3].  Builds the padded float4 / int4 device
 * buffers and the cached edge adjacency (replaces the per-call topology hash of dr.antialias). */
EHB_API int ehb_mesh_register(ehb_ctx_t ctx, const float* verts_host, int V, const int* faces_host, int F, int* mesh_id);
/* Replace the vertex positions of a registered mesh from a device buffer f32[V*3] (topology unchanged). */
EHB_API int ehb_mesh_update_verts(ehb_ctx_t ctx, int mesh_id, const float* verts_dev, int V, void* stream);
EHB_API int ehb_mesh_release(ehb_ctx_t ctx, int mesh_id);
EHB_API int ehb_mesh_info(ehb_ctx_t ctx, int mesh_id, int* V, int* F);

/* render_mask forward, one mesh, one pose.  anti_aliasing != 0: out_dev is f32[H*W] in [0,1];
 * anti_aliasing == 0: out_dev is u8[H*W] (0/1) = (z/w of nearest triangle > 0). */
EHB_API int ehb_render_mask_fwd(ehb_ctx_t ctx, int mesh_id, const float* mvp_dev, int H, int W, int anti_aliasing,
                        void* out_dev, void* stream);
/* render_mask backward (anti-aliased mask only): dy_dev f32[H*W] -> g_mvp_dev f64[16] (overwritten) and,
 * when g_pos_dev != NULL, the clip-space position gradient f32[V*4] (overwritten; x, y, w non-zero). */
EHB_API int ehb_render_mask_bwd(ehb_ctx_t ctx, int mesh_id, const float* mvp_dev, int H, int W, const float* dy_dev,
                        double* g_mvp_dev, float* g_pos_dev, void* stream);

/* RBSolver mask loop, fused: for b < B views and l < L links
 *     S_b = min(sum_l aa_mask(mesh_l, mvp[b,l]), 1)           -> masks_dev f32[B*H*W] (may be NULL)
 *     loss_b = sum_px (S_b - ref_b)^2                          -> loss_dev f64[B]      (needs ref_dev)
 *     g_mvp[b,l] = d( (1/B) sum_b loss_b ) / d mvp[b,l]        -> g_mvp_dev f64[B*L*16] (when do_bwd)
 * ref_dev f32[B*H*W] (NULL: masks only). */
EHB_API int ehb_render_views_fused(ehb_ctx_t ctx, const int* mesh_ids, int L, int B, const float* mvp_dev,
                           const float* ref_dev, int H, int W, int do_bwd, float* masks_dev, double* loss_dev,
                           double* g_mvp_dev, void* stream);

/* Same with the reference masks as bytes (non-zero = 1), the form EasyHeC's dataset holds them in before
 * `.float()` (easyhec/data/datasets/xarm_real.py:36,40): a quarter of the HBM and PCIe bytes. */
EHB_API int ehb_render_views_fused_u8(ehb_ctx_t ctx, const int* mesh_ids, int L, int B, const float* mvp_dev,
                              const uint8_t* ref_u8_dev, int H, int W, int do_bwd, float* masks_dev,
                              double* loss_dev, double* g_mvp_dev, void* stream);

/* Reference masks registered ONCE (they do not change during a solve; the reference trainer re-uploads them every step,
 * easyhec/trainer/rbsolver.py:31 `to_cuda(batch)`): binary masks (non-zero = 1; cv2.imread(path, 2) > 0,
 * easyhec/data/datasets/xarm_real.py:36,40) are bit-packed on the device together with their per-tile pixel counts, so a
 * fused step reads one bit per pixel of the tiles the robot touches and nothing of the others.
 *   ref: B*H*W values, image rows; dtype EHB_REF_U8 / EHB_REF_F32 (f32 values must be exactly 0 or 1);
 *   on_device != 0: `ref` is a device pointer.  Synchronous. */
#define EHB_REF_U8 0
#define EHB_REF_F32 1
EHB_API int ehb_ref_register(ehb_ctx_t ctx, const void* ref, int dtype, int on_device, int B, int H, int W, int* ref_id);
EHB_API int ehb_ref_release(ehb_ctx_t ctx, int ref_id);
/* ehb_render_views_fused against views [first_view, first_view + B) of a registered reference. */
EHB_API int ehb_render_views_fused_ref(ehb_ctx_t ctx, const int* mesh_ids, int L, int B, const float* mvp_dev, int ref_id,
                               int first_view, int H, int W, int do_bwd, float* masks_dev, double* loss_dev,
                               double* g_mvp_dev, void* stream);

/* N renders of the packed robot (all L links into one depth buffer, no anti-aliasing) -> out_dev u8[N*H*W]. */
EHB_API int ehb_render_binary_batch(ehb_ctx_t ctx, const int* mesh_ids, int L, int N, const float* mvp_dev, int H, int W,
                            uint8_t* out_dev, void* stream);
/* score[q] = sum_px unbiased_var_c(masks[q,c,px]), masks_dev u8[Q*C*n] -> score_dev f64[Q]. */
EHB_API int ehb_variance_score(ehb_ctx_t ctx, const uint8_t* masks_dev, int Q, int C, long long n, double* score_dev,
                       void* stream);
/* Fused space-exploration score: rasterizes the Q*C packed-robot renders (mvp_dev f32[Q*C*L*16]) into one depth plane each
 * and reduces the per-pixel variance over the C cameras of a candidate straight from those planes -- no mask is written
 * (easyhec/modeling/models/rb_solve/space_explorer.py:152-165). */
EHB_API int ehb_explore_scores(ehb_ctx_t ctx, const int* mesh_ids, int L, int Q, int C, const float* mvp_dev, int H, int W,
                       double* score_dev, void* stream);

/* Kinematic tree of a robot for the device-side forward kinematics (stands in for sapien / pinocchio,
 * easyhec/structures/sapien_kin.py:26-30).  Links in an order where parents precede children; per link: parent (-1 = root),
 * jtype (0 fixed, 1 revolute / continuous, 2 prismatic), qidx (index into qpos of its joint, -1 = none), joint value =
 * qpos[qidx] * mult + offs (mimic joints), unit axis f64[3], joint origin f64[16] (row-major 4x4, parent -> joint frame). */
EHB_API int ehb_robot_register(ehb_ctx_t ctx, int n_links, const int* parent, const int* jtype, const int* qidx,
                       const double* mult, const double* offs, const double* axis, const double* origin, int* robot_id);
/* mvp[q, c, l] = K_to_projection(K, H, W) @ diag(1,-1,-1,1) @ cams[c] @ FK(qpos[q])[sel_links[l]] for every candidate joint
 * configuration q (qpos_dev f64[Q*dof], device), camera pose c (cams_host f64[C*16] = Tc_c2b of each camera) and selected
 * link l -> mvp_dev f32[Q*C*L*16], the input of ehb_explore_scores (render_api.py:179-190 + nvdiffrast_renderer.py:33-37).
 * Synchronises the stream once (small host inputs). */
EHB_API int ehb_explore_fk_mvp(ehb_ctx_t ctx, int robot_id, const double* qpos_dev, int dof, int Q, const double* cams_host,
                       int C, const float* K_host, int H, int W, const int* sel_links, int L, float* mvp_dev, void* stream);

/* Host-buffer form of ehb_render_views_fused for callers without device pointers of their own: copies
 * mvp_host (pinned or pageable) to the device, runs the fused step on ref_dev, copies loss f64[B] and
 * g_mvp f64[B*L*16] back and synchronises the stream. */
EHB_API int ehb_solver_step_host(ehb_ctx_t ctx, const int* mesh_ids, int L, int B, const float* mvp_host,
                         const float* ref_dev, int H, int W, double* loss_host, double* g_mvp_host, void* stream);

/* Fully host-facing step: reference masks come from HOST memory as bytes (non-zero = 1; what the dataset holds
 * before `.float()`, easyhec/data/datasets/xarm_real.py:36) and are copied to the device inside the call, like
 * the reference trainer's per-step `to_cuda(batch)` (easyhec/trainer/rbsolver.py:31). */
EHB_API int ehb_solver_step_host_u8(ehb_ctx_t ctx, const int* mesh_ids, int L, int B, const float* mvp_host,
                            const uint8_t* ref_u8_host, int H, int W, double* loss_host, double* g_mvp_host,
                            void* stream);

/* The pose chain of RBSolver around the rasterizer, on the device (no host round trip, CUDA-graph capturable):
 *   ehb_pose_compose : dof f32[6] -> Tc_c2b = se3_exp_map(dof)^T (easyhec/utils/pytorch3d_se3.py:46-130, rb_solver.py:52)
 *                      -> mvp[b,l] = K_to_projection(K,H,W) @ diag(1,-1,-1,1) @ Tc_c2b @ link_poses[b,l]   f32[B*L*16]
 *   ehb_pose_backward: g_mvp f64[B*L*16], loss f64[B] -> out7 f32[7] = { grad_scale * d loss/d dof [6],
 *                      loss_scale * sum_b loss_b }   (one rank's share; all-reduce out7 across ranks when sharding views)
 *   ehb_adam_step    : torch.optim.Adam(lr, betas, eps, weight_decay as L2) on dof (easyhec/solver/build.py:12-29);
 *                      state f32[13] = { m[6], v[6], step }; hist_dev (optional) f32[hist_cap*6] records dof before
 *                      each update (RBSolver.history_ops, rb_solver.py:50-51). */
EHB_API int ehb_pose_compose(ehb_ctx_t ctx, const float* dof_dev, const float* K_dev, const float* link_poses_dev, int B,
                     int L, int H, int W, float* mvp_dev, void* stream);
EHB_API int ehb_pose_backward(ehb_ctx_t ctx, const float* dof_dev, const float* K_dev, const float* link_poses_dev,
                      const double* g_mvp_dev, const double* loss_dev, int B, int L, int H, int W, double grad_scale,
                      double loss_scale, float* out7_dev, void* stream);
EHB_API int ehb_adam_step(ehb_ctx_t ctx, float* dof_dev, const float* g7_dev, float* state_dev, float lr, float beta1,
                  float beta2, float eps, float weight_decay, float* hist_dev, int hist_cap, void* stream);

/* Asynchronous form of ehb_solver_step_host_u8 with four slots (slot = 0..3), so that the host<->device copies of one
 * step overlap the kernels of the others and the next copy is always queued behind the running one: _begin enqueues H2D(mvp, masks) -> fused step -> D2H(loss, g_mvp) on the slot's own
 * stream and returns; _end waits for that slot (the host buffers must stay valid and pinned until then).
 * _end returns EHB_E_OVERFLOW after growing the scratch if the step has to be submitted again. */
EHB_API int ehb_solver_step_begin_u8(ehb_ctx_t ctx, int slot, const int* mesh_ids, int L, int B, const float* mvp_host,
                             const uint8_t* ref_u8_host, int H, int W, double* loss_host, double* g_mvp_host);
EHB_API int ehb_solver_step_end(ehb_ctx_t ctx, int slot);
/* The same against a registered reference: per step only the matrices go up (B*L*64 bytes) and loss + gradient come down. */
EHB_API int ehb_solver_step_begin_ref(ehb_ctx_t ctx, int slot, const int* mesh_ids, int L, int B, const float* mvp_host,
                              int ref_id, int first_view, int H, int W, double* loss_host, double* g_mvp_host);

/* General asynchronous step on a slot (0..3) against registered reference masks: independent batches -- the views of
 * several solves, exploration rounds, ring slots of a benchmark -- run concurrently, each on its slot's stream and scratch.
 * The matrices come from the host (copied in) or are already on the device; optional outputs: the rendered masks (device),
 * loss / d loss/d mvp (host), and -- when dof_dev, K_dev, link_poses_dev are given -- the pose chain's out7 = { d loss/d dof
 * [6], mean loss } of the step (rb_solver.py:52-72 backward) on the device and / or the host, optionally followed by the
 * all-reduce of out7 over the ranks and the Adam update of a parameter vector.  ehb_solver_step_end(slot)
 * waits for the slot's last step and reports a scratch overflow; ehb_slots_fork / _join order all slot streams after / before
 * a caller's stream (for timing or for handing results on without a host wait). */
typedef struct ehb_step_io {
    const float* mvp_host;        /* f32[B*L*16], pinned host memory ... */
    const float* mvp_dev;         /* ... or on the device (one of the two) */
    float* masks_dev;             /* optional f32[B*H*W] */
    double* loss_host;            /* optional f64[B] */
    double* g_mvp_host;           /* optional f64[B*L*16] */
    const float* dof_dev;         /* optional pose chain: f32[6] */
    const float* K_dev;           /* f32[9] */
    const float* link_poses_dev;  /* f32[B*L*16] */
    float* out7_dev;              /* optional f32[7] */
    float* out7_host;             /* optional f32[7], pinned */
    float* adam_dof_dev;          /* optional Adam update behind the pose chain (trainer/rbsolver.py:29-43): the parameter f32[6] ... */
    float* adam_state_dev;        /* ... and its state f32[13] = { m[6], v[6], t } */
    double grad_scale, loss_scale;/* scales of out7 (0: 1 and 1 / B); a data-parallel caller passes 1 / world and 1 / (B * world) */
    float lr, weight_decay;       /* Adam (betas 0.9 / 0.999, eps 1e-8) */
    int exchange;                 /* != 0: out7 is summed over the ranks (ehb_comm_connect) between the pose chain and Adam, on the
                                     slot's own mailbox channel -- steps in flight on different slots never mix their messages, as
                                     long as every rank submits step k to the same slot */
    int pad;
} ehb_step_io_t;
EHB_API int ehb_step_begin(ehb_ctx_t ctx, int slot, const int* mesh_ids, int L, int B, int ref_id, int first_view, int H, int W,
                   const ehb_step_io_t* io);
/* the CUDA stream (cudaStream_t) of a slot, for callers that order their own work after a slot's step */
EHB_API int ehb_slot_stream(ehb_ctx_t ctx, int slot, void** stream);
EHB_API int ehb_slots_fork(ehb_ctx_t ctx, void* stream);
EHB_API int ehb_slots_join(ehb_ctx_t ctx, void* stream);

/* One-shot all-reduce (sum) of the 7 floats { d loss/d dof, loss } over NVLink peer memory, for view sharding across
 * the GPUs of one box -- the exchange DDP performs for the reference's 6-float parameter (easyhec/trainer/base.py:349).
 *   ehb_comm_local_handle: allocates this rank's mailbox, returns its 64-byte CUDA IPC handle
 *   ehb_comm_connect     : handles = world x 64 bytes (all ranks' handles, e.g. all-gathered with torch.distributed)
 *   ehb_allreduce7       : one tiny kernel on `stream`: writes its 7 floats into every peer's mailbox, waits for all
 *                          peers' values of the same step, sums in rank order (bit-identical on every rank).
 * Every rank must call ehb_allreduce7 the same number of times. */
EHB_API int ehb_comm_local_handle(ehb_ctx_t ctx, void* handle64);
EHB_API int ehb_comm_connect(ehb_ctx_t ctx, int rank, int world, const void* handles);
EHB_API int ehb_allreduce7(ehb_ctx_t ctx, float* g7_dev, void* stream);

/* The same exchange fused into the pose chain (one launch fewer per optimizer iteration): ehb_pose_backward_send computes
 * out7 like ehb_pose_backward and posts it to every peer's mailbox from the same kernel; ehb_adam_step_recv waits for all
 * ranks' messages of the step, adds them in rank order into g7_dev (the reduced values) and applies Adam.  Always used as
 * a pair, on the same stream, by every rank. */
EHB_API int ehb_pose_backward_send(ehb_ctx_t ctx, const float* dof_dev, const float* K_dev, const float* link_poses_dev,
                           const double* g_mvp_dev, const double* loss_dev, int B, int L, int H, int W, double grad_scale,
                           double loss_scale, float* out7_dev, void* stream);
EHB_API int ehb_adam_step_recv(ehb_ctx_t ctx, float* dof_dev, float* g7_dev, float* state_dev, float lr, float beta1,
                       float beta2, float eps, float weight_decay, float* hist_dev, int hist_cap, void* stream);
/* Adam update and, in the same launch, the matrices of the next iteration from the updated parameters (ehb_pose_compose's
 * arithmetic): mvp_dev f32[B*L*16].  recv != 0: the gradient is first summed over the ranks (as ehb_adam_step_recv). */
EHB_API int ehb_adam_step_compose(ehb_ctx_t ctx, float* dof_dev, float* g7_dev, float* state_dev, float lr, float beta1,
                                  float beta2, float eps, float weight_decay, float* hist_dev, int hist_cap, int recv,
                                  const float* K_dev, const float* link_poses_dev, int B, int L, int H, int W, float* mvp_dev,
                                  void* stream);

/* The tail of a solver iteration in ONE launch (trainer/rbsolver.py:29-43: loss.backward() through se3_exp_map, the DDP
 * all-reduce, optimizer.step()): ehb_pose_backward's out7, with exchange != 0 its all-reduce over the connected ranks, Adam on
 * adam_dof_dev (may be dof_dev itself) and, with mvp_next_dev != NULL, the matrices of the next iteration from the updated
 * parameters.  Same results as ehb_pose_backward[_send] followed by ehb_adam_step[_recv | _compose]. */
EHB_API int ehb_pose_backward_adam(ehb_ctx_t ctx, const float* dof_dev, const float* K_dev, const float* link_poses_dev,
                                   const double* g_mvp_dev, const double* loss_dev, int B, int L, int H, int W,
                                   double grad_scale, double loss_scale, float* out7_dev, int exchange, float* adam_dof_dev,
                                   float* state_dev, float lr, float beta1, float beta2, float eps, float weight_decay,
                                   float* hist_dev, int hist_cap, float* mvp_next_dev, void* stream);

/* Number of kernels this library has launched on the context since creation (for launch accounting). */
EHB_API long long ehb_launch_count(ehb_ctx_t ctx);

#ifdef __cplusplus
}
#endif
#endif
