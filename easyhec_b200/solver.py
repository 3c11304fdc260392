"""Pose optimisation loop on the device: RBSolver's Adam iteration as a CUDA graph.

The reference runs one Adam step per "epoch": zero_grad -> RBSolver.forward (B x L render_mask calls) ->
backward -> Adam (easyhec/trainer/rbsolver.py:29-43, easyhec/solver/build.py:12-29: lr 3e-3, weight decay 5e-4
as L2 on ``dof``).  Here one iteration is a fixed sequence of kernels

    [front, raster, raster_big, tiles] -> pose_adam: pose chain, (all-reduce of 7 floats), Adam, pose_compose of the next iteration

(five launches; with NCCL instead of the peer mailboxes the last one splits around the collective) captured into CUDA graphs
of 1 and 8 iterations and replayed; nothing returns to the host inside the loop.  When views are sharded
over ranks (``torch.distributed``), each rank renders its own views and the only exchange is the 7-float
all-reduce of (d loss/d dof, loss) -- the collective DDP performs for the reference (trainer/base.py:349).
"""
import numpy as np
import torch

from ._lib import Context, EhbError
from .se3 import dof_to_matrix, matrix_to_dof

__all__ = ["PoseSolver", "shard_views"]


def shard_views(n_views: int, rank: int, world: int):
    """Round-robin view indices of one rank (views are independent given dof, rb_solver.py:60-72)."""
    return list(range(rank, n_views, world))


class PoseSolver:
    def __init__(self, meshes, link_poses, K, masks_ref, init_Tc_c2b, H, W, lr=3e-3, weight_decay=5e-4,
                 betas=(0.9, 0.999), eps=1e-8, device=None, group=None, n_views_global=None, use_graph=True,
                 ctx=None, history=10000, peer_allreduce=True):
        """meshes: list of (verts, faces) / Mesh;  link_poses (B,L,4,4);  K (3,3);  masks_ref (B,H,W) bool/u8/f32
        -- this rank's views only;  n_views_global: total views over all ranks (defaults to B)."""
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.ctx = ctx or Context(self.device)
        self.H, self.W = int(H), int(W)
        self.mesh_ids = [self.ctx.register_mesh(*((m.vertices, m.faces) if hasattr(m, "vertices") else m))
                         for m in meshes]
        dev = self.device
        self.link_poses = torch.as_tensor(np.asarray(link_poses), dtype=torch.float32).to(dev).contiguous()
        self.K = torch.as_tensor(np.asarray(K), dtype=torch.float32).to(dev).contiguous()
        ref = torch.as_tensor(np.asarray(masks_ref)) if not isinstance(masks_ref, torch.Tensor) else masks_ref
        if ref.dtype == torch.bool:
            ref = ref.to(torch.uint8)
        if ref.dtype not in (torch.uint8, torch.float32):
            ref = ref.float()
        # the reference masks do not change during a solve: registered once (bit-packed, per-tile counts); soft (non-binary)
        # masks cannot be packed and stay a plain f32 tensor
        try:
            self.ref = self.ctx.register_ref(ref)
        except EhbError:
            self.ref = ref.to(dev).float().contiguous()
        self.B, self.L = self.link_poses.shape[0], self.link_poses.shape[1]
        self.B_global = int(n_views_global or self.B)
        self.group = group
        self.world = torch.distributed.get_world_size(group) if (group is not None or (
            torch.distributed.is_available() and torch.distributed.is_initialized())) else 1
        self.lr, self.wd, self.betas, self.eps = float(lr), float(weight_decay), betas, float(eps)
        self.dof = matrix_to_dof(torch.as_tensor(np.asarray(init_Tc_c2b), dtype=torch.float32)).to(dev).contiguous()
        self.state = torch.zeros(13, dtype=torch.float32, device=dev)
        self.hist = torch.zeros((history, 6), dtype=torch.float32, device=dev) if history else None
        self.mvp = torch.empty((self.B, self.L, 4, 4), dtype=torch.float32, device=dev)
        self.loss_b = torch.empty((self.B,), dtype=torch.float64, device=dev)
        self.g_mvp = torch.empty((self.B, self.L, 4, 4), dtype=torch.float64, device=dev)
        self.g7 = torch.zeros(7, dtype=torch.float32, device=dev)
        self.masks = None
        F = sum(self.ctx.mesh_info(i)[1] for i in self.mesh_ids)
        self.ctx.reserve(self.B, self.L, F, self.H, self.W)
        self.use_graph = use_graph
        self._peer = False
        if self.world > 1 and peer_allreduce:
            try:   # NVLink peer-mailbox all-reduce instead of a NCCL launch on the critical path
                self.ctx.comm_connect(group)
                self._peer = True
            except Exception:   # noqa: BLE001 -- e.g. GPUs without peer access: NCCL does the same job
                self._peer = False
            ok = torch.tensor([1.0 if self._peer else 0.0], device=dev)
            torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN, group=group)
            self._peer = ok.item() >= 1.0
        self._graph, self._graph_n = None, None
        self.unroll = 8              # iterations per captured graph
        self._mvp_valid = False      # self.mvp == compose(self.dof)
        import os
        self._fuse_compose = not os.environ.get("EHB_SOLVER_NOFUSE")
        self.iterations = 0

    # one iteration, enqueued on the current stream
    def _iteration(self):
        c = self.ctx
        if not self._mvp_valid:      # (outside the captured graph: the first iteration, or after a roll-back)
            c.pose_compose(self.dof, self.K, self.link_poses, self.H, self.W, out=self.mvp)
            self._mvp_valid = True
        c.render_views_fused(self.mesh_ids, self.mvp, self.ref, self.H, self.W, backward=True,
                             out=(self.masks, self.loss_b, self.g_mvp))
        # fused kernel scaled by 1/B_local; rescale so that the sum over ranks is the global mean's gradient
        fusedx = self.world > 1 and self._peer      # the 7-float exchange rides inside pose_backward (send) and adam (recv)
        gs, ls = self.B / self.B_global, 1.0 / self.B_global
        if self._fuse_compose and (self.world == 1 or self._peer):
            # the tail of the iteration in one launch: pose chain (+ send), (recv +) Adam, the next iteration's matrices
            c.pose_backward_adam(self.dof, self.K, self.link_poses, self.g_mvp, self.loss_b, self.H, self.W, self.state, self.lr,
                                 self.betas, self.eps, self.wd, grad_scale=gs, loss_scale=ls, out=self.g7, exchange=fusedx,
                                 hist=self.hist, mvp_next=self.mvp)
            return
        c.pose_backward(self.dof, self.K, self.link_poses, self.g_mvp, self.loss_b, self.H, self.W,
                        grad_scale=gs, loss_scale=ls, out=self.g7, send=fusedx)
        if self.world > 1 and not self._peer:
            torch.distributed.all_reduce(self.g7, group=self.group)
        # Adam, and in the same launch the matrices of the next iteration from the updated parameters
        if self._fuse_compose:
            c.adam_step(self.dof, self.g7, self.state, self.lr, self.betas, self.eps, self.wd, hist=self.hist, recv=fusedx,
                        compose=(self.K, self.link_poses, self.H, self.W, self.mvp))
        else:
            c.adam_step(self.dof, self.g7, self.state, self.lr, self.betas, self.eps, self.wd, hist=self.hist, recv=fusedx)
            c.pose_compose(self.dof, self.K, self.link_poses, self.H, self.W, out=self.mvp)

    def _overflowed(self) -> bool:
        """Synchronises; True when a launch since the last look overflowed a scratch pool on ANY rank (the decision to
        redo iterations must be the same everywhere, or the ranks' exchanges would fall out of step)."""
        # one stream synchronisation per chunk, then the flags from host-visible memory (ehb_ctx_poll); the full status
        # read-out -- a device synchronisation and two copies of every counter block -- only when a flag is up
        torch.cuda.current_stream(self.device).synchronize()
        bad = bool(self.ctx.poll() & 1)
        if bad or self.ctx.poll():
            self.ctx.status()
        if self.world > 1:
            t = torch.tensor([1.0 if bad else 0.0], device=self.device)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX, group=self.group)
            bad = t.item() > 0
        return bad

    def _capture(self):
        """One graph of a single iteration and one of `unroll` iterations back to back: inside a graph the kernels of
        consecutive iterations are chained (programmatic dependent launch); between graph launches the stream idles for a few
        microseconds -- 4 % of an iteration at 640x480."""
        if not self._mvp_valid:      # the captured iteration starts from valid matrices; it never composes them first
            self.ctx.pose_compose(self.dof, self.K, self.link_poses, self.H, self.W, out=self.mvp)
            self._mvp_valid = True
        torch.cuda.synchronize(self.device)
        graphs = []
        for reps in (1, self.unroll):
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream(self.device)
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s):
                with torch.cuda.graph(g, stream=s):     # capture does not execute: dof / state are untouched
                    for _ in range(reps):
                        self._iteration()
            torch.cuda.current_stream(self.device).wait_stream(s)
            graphs.append(g)
        self._graph, self._graph_n = graphs

    def step(self, n: int = 1, check_every: int = 128):
        """Run n Adam iterations.  Iterations are replayed in chunks of `check_every` with no host synchronisation inside
        a chunk; after each chunk the sticky status flags are read once.  A scratch-pool overflow (the silhouette grew
        beyond what was reserved) restores the chunk's starting state, grows the scratch, re-captures the graph and
        runs the chunk again -- on every rank alike."""
        done = 0
        while done < n:
            m = min(check_every, n - done) if self.iterations > 0 else 1     # the very first iteration sizes the scratch
            snap = (self.dof.clone(), self.state.clone())
            for _attempt in range(12):
                if self.use_graph and self.iterations > 0 and self._graph is None:
                    self._capture()
                if self._graph is not None:
                    for _ in range(m // self.unroll):
                        self._graph_n.replay()
                    for _ in range(m % self.unroll):
                        self._graph.replay()
                else:
                    for _ in range(m):
                        self._iteration()
                if not self._overflowed():
                    break
                self.dof.copy_(snap[0]); self.state.copy_(snap[1])
                self._mvp_valid = False
                self.ctx.grow_scratch()
                self._graph = None                  # scratch buffers move when they grow: the captured pointers are stale
                F = sum(self.ctx.mesh_info(i)[1] for i in self.mesh_ids)
                self.ctx.reserve(self.B, self.L, F, self.H, self.W)
            else:
                raise EhbError("the rasterizer's scratch pools kept overflowing")
            self.iterations += m
            done += m

    @property
    def loss(self) -> torch.Tensor:
        """Mean mask loss of the last iteration (device scalar, before that iteration's update)."""
        return self.g7[6]

    def Tc_c2b(self) -> torch.Tensor:
        return dof_to_matrix(self.dof.detach().cpu())

    def history_ops(self) -> torch.Tensor:
        n = int(self.state[12].item())
        return self.hist[:n].clone() if self.hist is not None else torch.zeros(0, 6)

    def pose_error(self, gt_Tc_c2b):
        """(translation error in metres, rotation error in degrees) against a ground-truth pose."""
        T = self.Tc_c2b().double().numpy()
        G = np.asarray(gt_Tc_c2b, dtype=np.float64)
        # ||R_T - R_G||_F = 2 sqrt(2) sin(angle / 2): well conditioned near zero, unlike arccos((trace - 1) / 2)
        ang = np.degrees(2.0 * np.arcsin(min(1.0, np.linalg.norm(T[:3, :3] - G[:3, :3]) / (2.0 * np.sqrt(2.0)))))
        return float(np.linalg.norm(T[:3, 3] - G[:3, 3])), float(ang)
