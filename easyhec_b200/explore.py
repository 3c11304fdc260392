"""Space-exploration scoring on the device.

The reference scores every candidate joint configuration by rendering the whole robot from each camera pose sampled
so far and summing the per-pixel variance of the binary masks over those cameras
(easyhec/modeling/models/rb_solve/space_explorer.py:152-165, easyhec/utils/render_api.py:70-96,179-192): one URDF
forward kinematics + mesh re-pack + upload + render per (candidate, camera), ten thousand sequential renders per round.
Here the forward kinematics of ALL candidates is one batched torch evaluation (``URDFKinematics``, on the device when
the inputs are), the model-view-projection matrices of every (candidate, camera, link) are one broadcast product, and
``ehb_explore_scores`` renders and scores them without a mask ever leaving the GPU.  Collision / reachability filters
of the reference (planner, max-distance constraint) stay on the host and enter as ``valid``: rejected candidates score
0, exactly like the reference's ``variances.append(0)``.
"""
import numpy as np
import torch

from .projection import K_to_projection, opencv2gl

__all__ = ["candidate_mvps", "score_candidates", "select_next_qpos", "shard_candidates"]


def candidate_mvps(kin, links, qposes, cam_poses, K, H, W, pad_left=0, pad_right=0, dtype=torch.float32):
    """(Q, dof') candidate joint positions x (C, 4, 4) camera poses (``Tc_c2b`` of each camera: base -> camera, i.e. the
    ``np.linalg.inv(cam_pose)`` the reference hands to the renderer, space_explorer.py:153-155) -> mvp (Q, C, L, 4, 4).

    mvp[q, c, l] = K_to_projection(K) @ diag(1,-1,-1,1) @ Tc_c2b[c] @ FK(qpos_q)[links[l]]  (nvdiffrast_renderer.py:33-37).
    ``pad_left`` / ``pad_right`` zeros are added around every qpos like ``qpos_choices_pad_left/right``
    (space_explorer.py:103)."""
    q = torch.as_tensor(np.asarray(qposes) if not isinstance(qposes, torch.Tensor) else qposes)
    if not q.is_floating_point():
        q = q.double()
    if q.ndim == 1:
        q = q[None]
    if pad_left or pad_right:
        q = torch.cat([q.new_zeros(q.shape[0], pad_left), q, q.new_zeros(q.shape[0], pad_right)], 1)
    link_poses = kin.forward(q, links=links)                                   # (Q, L, 4, 4), same device as q
    cams = torch.as_tensor(np.asarray(cam_poses) if not isinstance(cam_poses, torch.Tensor) else cam_poses)
    cams = cams.to(device=link_poses.device, dtype=link_poses.dtype)
    if cams.ndim == 2:
        cams = cams[None]
    Kt = torch.as_tensor(np.asarray(K) if not isinstance(K, torch.Tensor) else K).float()
    P = (K_to_projection(Kt, H, W) @ opencv2gl()).to(device=link_poses.device, dtype=link_poses.dtype)
    mvp = P[None, None, None] @ (cams[None, :, None] @ link_poses[:, None])    # (Q, C, L, 4, 4)
    return mvp.to(dtype).contiguous()


def shard_candidates(n: int, rank: int, world: int):
    """Block partition of n candidates over `world` ranks: (first, count) of one rank (candidates are independent,
    space_explorer.py:99)."""
    base, extra = divmod(n, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def score_candidates(ctx, mesh_ids, kin, links, qposes, cam_poses, K, H, W, valid=None, pad_left=0, pad_right=0,
                     device_fk=True, group=None, robot=None, check=True):
    """Variance score of every candidate (space_explorer.py:163-164: ``torch.var(masks, dim=0).sum()``, unbiased) as a
    (Q,) float64 tensor on the context's device; candidates with ``valid[q] == False`` score 0.

    device_fk: forward kinematics and matrix composition run on the GPU (``ehb_explore_fk_mvp``); otherwise
    ``candidate_mvps`` composes them with torch.  group / an initialised ``torch.distributed``: every rank scores its block of
    the candidates and ONE all-gather of the scores follows -- the only exchange of this path."""
    q = torch.as_tensor(np.asarray(qposes) if not isinstance(qposes, torch.Tensor) else qposes).double()
    if q.ndim == 1:
        q = q[None]
    if pad_left or pad_right:
        q = torch.cat([q.new_zeros(q.shape[0], pad_left), q, q.new_zeros(q.shape[0], pad_right)], 1)
    Q = q.shape[0]
    import torch.distributed as dist
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    first, count = shard_candidates(Q, rank, world)
    mine = q[first:first + count]
    if count == 0:
        local = torch.zeros(0, dtype=torch.float64, device=ctx.device)
    else:
        if device_fk:
            robot = robot or ctx.register_robot(kin)
            mvp = ctx.explore_fk_mvp(robot, mine.to(ctx.device).contiguous(), cam_poses, K, H, W, links)
        else:
            mvp = candidate_mvps(kin, links, mine, cam_poses, K, H, W).to(ctx.device)
        local = ctx.explore_scores(mesh_ids, mvp, H, W)
        if check and hasattr(ctx, "status"):
            # the scores are about to be read by the host anyway: look at the sticky flags once per batch; an overflowed
            # scratch pool (incomplete renders) is grown and the batch scored again
            for _ in range(12):
                flags, _ = ctx.status()
                if not flags & 1:
                    break
                ctx.grow_scratch()
                local = ctx.explore_scores(mesh_ids, mvp, H, W)
    if world > 1:
        per = -(-Q // world)                               # equal-sized slots for the all-gather, trimmed afterwards
        buf = torch.zeros(per, dtype=torch.float64, device=ctx.device)
        buf[:count] = local
        out = torch.empty(world * per, dtype=torch.float64, device=ctx.device)
        dist.all_gather_into_tensor(out, buf, group=group)
        scores = torch.cat([out[r * per:r * per + shard_candidates(Q, r, world)[1]] for r in range(world)])
    else:
        scores = local
    if valid is not None:
        scores = torch.where(torch.as_tensor(np.asarray(valid), dtype=torch.bool, device=scores.device), scores,
                             torch.zeros_like(scores))
    return scores


def select_next_qpos(scores, history=None):
    """Index of the best candidate that has not been selected before (space_explorer.py:171-180 keeps a history of the
    chosen candidates), and its score."""
    s = scores.detach().clone()
    if history:
        s[torch.as_tensor(list(history), dtype=torch.long, device=s.device)] = -1.0
    i = int(torch.argmax(s).item())
    return i, float(s[i].item())
