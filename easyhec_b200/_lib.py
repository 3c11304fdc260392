"""ctypes binding of libehb.so (the C ABI in include/easyhec_b200.h).

PyTorch is plumbing here: it owns device memory and the current stream; every compute call goes through
the C ABI with raw pointers.  There is no CPU or eager fallback -- if the library is missing or there is
no CUDA device the calls raise.
"""
import ctypes as C
import os

import numpy as np
import torch

__all__ = ["EhbError", "lib", "Context", "RefMasks", "library_path"]

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("EHB_LIB") or os.path.join(_HERE, "libehb.so")
_lib = None

EHB_FLAG_PAIR_OVERFLOW = 1
EHB_FLAG_NEEDS_CLIP = 2
EHB_FLAG_QUEUES_FULL = 4


class EhbError(RuntimeError):
    pass


def library_path() -> str:
    return _SO


_PROTOS = {
    "ehb_version": (C.c_int, []),
    "ehb_last_error": (C.c_char_p, []),
    "ehb_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "ehb_ctx_destroy": (C.c_int, [C.c_void_p]),
    "ehb_ctx_reserve": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ehb_ctx_set_fill_rule": (C.c_int, [C.c_void_p, C.c_int]),
    "ehb_ctx_set_pool_budget": (C.c_int, [C.c_void_p, C.c_double]),
    "ehb_ctx_set_pipelines": (C.c_int, [C.c_void_p, C.c_int]),
    "ehb_ctx_grow_scratch": (C.c_int, [C.c_void_p]),
    "ehb_ctx_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "ehb_ctx_kernel_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "ehb_ctx_kernel_times_peek": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "ehb_ctx_debug_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_ulonglong), C.c_int]),
    "ehb_ctx_debug_buffer": (C.c_int, [C.c_void_p, C.POINTER(C.c_ulonglong), C.c_int]),
    "ehb_ctx_debug_marks": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint)]),
    "ehb_ctx_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint), C.POINTER(C.c_longlong)]),
    "ehb_ctx_poll": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint)]),
    "ehb_mesh_register": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "ehb_mesh_update_verts": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "ehb_mesh_release": (C.c_int, [C.c_void_p, C.c_int]),
    "ehb_mesh_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ehb_render_mask_fwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p]),
    "ehb_render_mask_bwd": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "ehb_render_views_fused": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ehb_render_views_fused_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                            C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ehb_ref_register": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "ehb_ref_release": (C.c_int, [C.c_void_p, C.c_int]),
    "ehb_render_views_fused_ref": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                             C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ehb_render_binary_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p]),
    "ehb_variance_score": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p]),
    "ehb_explore_scores": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p]),
    "ehb_robot_register": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.POINTER(C.c_int)]),
    "ehb_explore_fk_mvp": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                     C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "ehb_solver_step_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ehb_solver_step_host_u8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ehb_pose_compose": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p]),
    "ehb_pose_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
    "ehb_pose_backward_send": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
    "ehb_adam_step_recv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float,
                                     C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_void_p]),
    "ehb_adam_step_compose": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float,
                                        C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_void_p, C.c_void_p]),
    "ehb_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float,
                                C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_void_p]),
    "ehb_pose_backward_adam": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int,
                                         C.c_void_p, C.c_void_p]),
    "ehb_solver_step_begin_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ehb_solver_step_end": (C.c_int, [C.c_void_p, C.c_int]),
    "ehb_solver_step_begin_ref": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                            C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ehb_step_begin": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ehb_slot_stream": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "ehb_slots_fork": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ehb_slots_join": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ehb_comm_local_handle": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ehb_comm_connect": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "ehb_allreduce7": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ehb_launch_count": (C.c_longlong, [C.c_void_p]),
}


def lib():
    """Load libehb.so (built by `make -C easyhec_b200/csrc` / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise EhbError("libehb.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "or `make -C easyhec_b200/csrc`); there is no fallback path")
        l = C.CDLL(_SO)
        for name, (res, args) in _PROTOS.items():
            f = getattr(l, name)
            f.restype = res
            f.argtypes = args
        _lib = l
    return _lib


def _check(rc: int):
    if rc != 0:
        msg = lib().ehb_last_error()
        raise EhbError("libehb error %d: %s" % (rc, msg.decode("utf-8", "replace") if msg else "?"))


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _dev_check(t, dtype, device, name):
    if t is None:
        return
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.device != device:
        raise EhbError("%s must be a CUDA tensor on %s" % (name, device))
    if t.dtype != dtype:
        raise EhbError("%s must have dtype %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise EhbError("%s must be contiguous" % name)


class StepIO(C.Structure):
    """ehb_step_io_t (include/easyhec_b200.h)"""
    _fields_ = ([(n, C.c_void_p) for n in ("mvp_host", "mvp_dev", "masks_dev", "loss_host", "g_mvp_host", "dof_dev", "K_dev",
                                           "link_poses_dev", "out7_dev", "out7_host", "adam_dof_dev", "adam_state_dev")] +
                [("grad_scale", C.c_double), ("loss_scale", C.c_double), ("lr", C.c_float), ("weight_decay", C.c_float),
                 ("exchange", C.c_int), ("pad", C.c_int)])


class RefMasks:
    """Reference masks registered once with a context (`Context.register_ref`): bit-packed on the device.  `first` / `n`
    select a run of its views (`ref[2:6]` = views 2..5), so one registration serves sharded or sliced calls."""

    def __init__(self, ctx, ref_id, n, H, W, first=0):
        self.ctx, self.ref_id, self.n, self.H, self.W, self.first = ctx, ref_id, int(n), int(H), int(W), int(first)

    def __len__(self):
        return self.n

    def __getitem__(self, sl):
        if not isinstance(sl, slice) or sl.step not in (None, 1):
            raise EhbError("registered reference masks can only be sliced into contiguous runs of views")
        a, b, _ = sl.indices(self.n)
        return RefMasks(self.ctx, self.ref_id, max(b - a, 0), self.H, self.W, self.first + a)

    def release(self):
        self.ctx.release_ref(self)


class Context:
    """One rasterizer context per device (replaces dr.RasterizeCudaContext, nvdiffrast_renderer.py:23)."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise EhbError("no CUDA device: easyhec_b200 has no CPU path")
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        torch.cuda.init()
        with torch.cuda.device(device):
            torch.zeros(1, device=device)  # make sure the primary context exists
        h = C.c_void_p()
        _check(lib().ehb_ctx_create(device.index, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib().ehb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- meshes ------------------------------------------------------------------------------------------
    def register_mesh(self, verts, faces) -> int:
        v = np.ascontiguousarray(verts.detach().cpu().numpy() if isinstance(verts, torch.Tensor) else verts,
                                 dtype=np.float32).reshape(-1, 3)
        f = np.ascontiguousarray(faces.detach().cpu().numpy() if isinstance(faces, torch.Tensor) else faces,
                                 dtype=np.int32).reshape(-1, 3)
        mid = C.c_int(-1)
        _check(lib().ehb_mesh_register(self._h, v.ctypes.data_as(C.c_void_p), len(v), f.ctypes.data_as(C.c_void_p),
                                       len(f), C.byref(mid)))
        return mid.value

    def update_verts(self, mesh_id: int, verts: torch.Tensor):
        _dev_check(verts, torch.float32, self.device, "verts")
        _check(lib().ehb_mesh_update_verts(self._h, mesh_id, _ptr(verts), verts.shape[0], _stream(self.device)))

    def release_mesh(self, mesh_id: int):
        _check(lib().ehb_mesh_release(self._h, mesh_id))

    def mesh_info(self, mesh_id: int):
        V, F = C.c_int(), C.c_int()
        _check(lib().ehb_mesh_info(self._h, mesh_id, C.byref(V), C.byref(F)))
        return V.value, F.value

    # -- reference masks -----------------------------------------------------------------------------------
    def register_ref(self, masks) -> RefMasks:
        """masks (B,H,W) bool / uint8 / float32 (values 0 or 1), numpy or torch, host or device -> RefMasks handle."""
        if isinstance(masks, np.ndarray):
            masks = torch.from_numpy(np.ascontiguousarray(masks))
        if masks.dtype == torch.bool:
            masks = masks.to(torch.uint8)
        if masks.dtype not in (torch.uint8, torch.float32):
            masks = masks.to(torch.float32)
        if masks.dim() != 3:
            raise EhbError("reference masks must have shape (B, H, W)")
        masks = masks.contiguous()
        if masks.is_cuda and masks.device != self.device:
            raise EhbError("reference masks live on %s, the context on %s" % (masks.device, self.device))
        B, H, W = masks.shape
        rid = C.c_int(-1)
        _check(lib().ehb_ref_register(self._h, _ptr(masks), 1 if masks.dtype == torch.float32 else 0, int(masks.is_cuda),
                                      B, H, W, C.byref(rid)))
        return RefMasks(self, rid.value, B, H, W)

    def release_ref(self, ref: RefMasks):
        _check(lib().ehb_ref_release(self._h, ref.ref_id))

    # -- control -----------------------------------------------------------------------------------------
    def reserve(self, n_items, n_links, max_faces, H, W):
        _check(lib().ehb_ctx_reserve(self._h, n_items, n_links, max_faces, H, W))

    def set_fill_rule(self, rule: int):
        _check(lib().ehb_ctx_set_fill_rule(self._h, rule))

    def set_pool_budget(self, nbytes: float):
        _check(lib().ehb_ctx_set_pool_budget(self._h, float(nbytes)))

    def set_pipelines(self, n: int):
        _check(lib().ehb_ctx_set_pipelines(self._h, int(n)))

    def grow_scratch(self):
        _check(lib().ehb_ctx_grow_scratch(self._h))

    def status(self):
        """Synchronises; returns (flags, n_need_clip) and clears them."""
        fl, nc = C.c_uint(), C.c_longlong()
        _check(lib().ehb_ctx_status(self._h, C.byref(fl), C.byref(nc)))
        return fl.value, nc.value

    def poll(self) -> int:
        """The sticky flags without synchronising (may lag behind work that is still running); cleared by status()."""
        fl = C.c_uint()
        _check(lib().ehb_ctx_poll(self._h, C.byref(fl)))
        return fl.value

    def check(self, what="launch"):
        """Raise if an earlier launch overflowed a scratch pool (its results are incomplete).  Non-synchronising."""
        if self.poll() & EHB_FLAG_PAIR_OVERFLOW:
            raise EhbError("a scratch pool of the rasterizer overflowed during an earlier %s: results of that launch are "
                           "incomplete (call status(), grow_scratch() and run it again)" % what)

    def profile(self, enable: bool):
        _check(lib().ehb_ctx_profile(self._h, int(bool(enable))))

    def kernel_times(self, peek=False):
        """-> ({front (table | vertices | batch lists | clear | tile lists), raster (+ raster_big), tiles: summed ms}, passes);
        synchronises.  peek=True keeps the recorded passes (passes captured in a CUDA graph: read after every replay)."""
        ms = (C.c_double * 4)()
        n = C.c_longlong()
        _check((lib().ehb_ctx_kernel_times_peek if peek else lib().ehb_ctx_kernel_times)(self._h, ms, C.byref(n)))
        return {"front": ms[0] + ms[1], "raster": ms[2], "tiles": ms[3]}, n.value

    def debug_counters(self, reset=True):
        out = (C.c_ulonglong * 16)()
        _check(lib().ehb_ctx_debug_counters(self._h, out, int(reset)))
        return list(out)

    def debug_buffer(self, n_words=0):
        out = (C.c_ulonglong * max(n_words, 1))()
        _check(lib().ehb_ctx_debug_buffer(self._h, out, n_words))
        return list(out)[:n_words]

    def debug_marks(self):
        out = (C.c_uint * 16)()
        _check(lib().ehb_ctx_debug_marks(self._h, out))
        return list(out)

    def launch_count(self) -> int:
        return int(lib().ehb_launch_count(self._h))

    # -- compute -----------------------------------------------------------------------------------------
    def render_mask_fwd(self, mesh_id, mvp, H, W, anti_aliasing=True):
        _dev_check(mvp, torch.float32, self.device, "mvp")
        out = torch.empty((H, W), dtype=torch.float32 if anti_aliasing else torch.uint8, device=self.device)
        _check(lib().ehb_render_mask_fwd(self._h, mesh_id, _ptr(mvp), H, W, int(bool(anti_aliasing)), _ptr(out),
                                         _stream(self.device)))
        return out

    def render_mask_bwd(self, mesh_id, mvp, H, W, dy, want_gpos=False):
        _dev_check(mvp, torch.float32, self.device, "mvp")
        _dev_check(dy, torch.float32, self.device, "dy")
        g_mvp = torch.empty((4, 4), dtype=torch.float64, device=self.device)
        g_pos = None
        if want_gpos:
            V, _ = self.mesh_info(mesh_id)
            g_pos = torch.empty((V, 4), dtype=torch.float32, device=self.device)
        _check(lib().ehb_render_mask_bwd(self._h, mesh_id, _ptr(mvp), H, W, _ptr(dy), _ptr(g_mvp), _ptr(g_pos),
                                         _stream(self.device)))
        return g_mvp, g_pos

    def render_views_fused(self, mesh_ids, mvp, ref, H, W, backward=True, want_masks=True, out=None):
        """mvp (B,L,4,4) f32, ref (B,H,W) f32 or u8, a RefMasks handle, or None -> masks (B,H,W) f32 | None, loss (B,) f64,
        g_mvp (B,L,4,4) f64 | None.  `out` = (masks, loss, g_mvp) pre-allocated tensors to reuse."""
        _dev_check(mvp, torch.float32, self.device, "mvp")
        B, L = mvp.shape[0], mvp.shape[1]
        ids = (C.c_int * L)(*mesh_ids)
        if out is not None:
            masks, loss, g_mvp = out
        else:
            masks = torch.empty((B, H, W), dtype=torch.float32, device=self.device) if want_masks else None
            loss = torch.empty((B,), dtype=torch.float64, device=self.device) if ref is not None else None
            g_mvp = torch.empty((B, L, 4, 4), dtype=torch.float64, device=self.device) if backward else None
        if isinstance(ref, RefMasks):
            if ref.ctx is not self or len(ref) != B or (ref.H, ref.W) != (H, W):
                raise EhbError("registered reference: %d views of %dx%d on another context or of another shape" % (len(ref), ref.H, ref.W))
            _check(lib().ehb_render_views_fused_ref(self._h, ids, L, B, _ptr(mvp), ref.ref_id, ref.first, H, W,
                                                    int(bool(backward)), _ptr(masks), _ptr(loss), _ptr(g_mvp),
                                                    _stream(self.device)))
        elif ref is not None and ref.dtype == torch.uint8:
            _dev_check(ref, torch.uint8, self.device, "ref")
            _check(lib().ehb_render_views_fused_u8(self._h, ids, L, B, _ptr(mvp), _ptr(ref), H, W, int(bool(backward)),
                                                   _ptr(masks), _ptr(loss), _ptr(g_mvp), _stream(self.device)))
        else:
            _dev_check(ref, torch.float32, self.device, "ref")
            _check(lib().ehb_render_views_fused(self._h, ids, L, B, _ptr(mvp), _ptr(ref), H, W, int(bool(backward)),
                                                _ptr(masks), _ptr(loss), _ptr(g_mvp), _stream(self.device)))
        return masks, loss, g_mvp

    def render_binary_batch(self, mesh_ids, mvp, H, W, out=None):
        """mvp (N,L,4,4) -> u8 (N,H,W): packed robot, one depth buffer per render, no antialiasing."""
        _dev_check(mvp, torch.float32, self.device, "mvp")
        N, L = mvp.shape[0], mvp.shape[1]
        ids = (C.c_int * L)(*mesh_ids)
        if out is None:
            out = torch.empty((N, H, W), dtype=torch.uint8, device=self.device)
        _check(lib().ehb_render_binary_batch(self._h, ids, L, N, _ptr(mvp), H, W, _ptr(out), _stream(self.device)))
        return out

    def variance_score(self, masks):
        """masks (Q,C,H,W) u8/bool -> (Q,) f64 = sum_px var_c (unbiased)."""
        if masks.dtype == torch.bool:
            masks = masks.view(torch.uint8)
        _dev_check(masks, torch.uint8, self.device, "masks")
        Q, Cn = masks.shape[0], masks.shape[1]
        n = masks[0, 0].numel() if Q else 0
        score = torch.empty((Q,), dtype=torch.float64, device=self.device)
        _check(lib().ehb_variance_score(self._h, _ptr(masks), Q, Cn, n, _ptr(score), _stream(self.device)))
        return score

    def explore_scores(self, mesh_ids, mvp, H, W):
        """mvp (Q,C,L,4,4) -> (Q,) f64 variance scores (space_explorer.py:152-165), masks never leave the GPU."""
        _dev_check(mvp, torch.float32, self.device, "mvp")
        Q, Cn, L = mvp.shape[0], mvp.shape[1], mvp.shape[2]
        ids = (C.c_int * L)(*mesh_ids)
        score = torch.empty((Q,), dtype=torch.float64, device=self.device)
        _check(lib().ehb_explore_scores(self._h, ids, L, Q, Cn, _ptr(mvp), H, W, _ptr(score), _stream(self.device)))
        return score

    # -- space exploration: forward kinematics on the device ---------------------------------------------------------------
    def register_robot(self, kin):
        """kin: URDFKinematics -> handle for explore_fk_mvp.  The kinematic tree goes to the device once (links in an order
        where parents precede children; the handle keeps the map from the URDF's link order)."""
        order, parent_of = [kin.root_link], {kin.root_link: None}
        pending = [j for j in kin.joints]
        while pending:
            rest = [j for j in pending if j["parent"] not in parent_of]
            for j in pending:
                if j["parent"] in parent_of and j["child"] not in parent_of:
                    parent_of[j["child"]] = j
                    order.append(j["child"])
            if len(rest) == len(pending):
                break
            pending = rest
        idx = {n: i for i, n in enumerate(order)}
        qi = {j["name"]: i for i, j in enumerate(kin.movable)}
        n = len(order)
        parent = np.full(n, -1, np.int32); jtype = np.zeros(n, np.int32); qidx = np.full(n, -1, np.int32)
        mult = np.ones(n); offs = np.zeros(n); axis = np.zeros((n, 3)); origin = np.tile(np.eye(4), (n, 1, 1))
        for name in order[1:]:
            j, i = parent_of[name], idx[name]
            parent[i] = idx[j["parent"]]
            origin[i] = j["origin"]
            axis[i] = j["axis"]
            if j["type"] in ("revolute", "continuous", "prismatic"):
                jtype[i] = 2 if j["type"] == "prismatic" else 1
                if j["mimic"] is not None and j["mimic"][0] in qi:
                    qidx[i], mult[i], offs[i] = qi[j["mimic"][0]], j["mimic"][1], j["mimic"][2]
                elif j["name"] in qi:
                    qidx[i] = qi[j["name"]]
                else:
                    jtype[i] = 0
        rid = C.c_int(-1)
        arrs = [np.ascontiguousarray(a) for a in (parent, jtype, qidx, mult, offs, axis, origin)]
        _check(lib().ehb_robot_register(self._h, n, *[a.ctypes.data_as(C.c_void_p) for a in arrs], C.byref(rid)))
        return {"id": rid.value, "tree_index": {name: idx[name] for name in kin.link_names if name in idx},
                "link_names": list(kin.link_names), "dof": kin.dof}

    def explore_fk_mvp(self, robot, qpos, cam_poses, K, H, W, links):
        """qpos (Q, <= dof) float64 CUDA tensor, cam_poses (C,4,4), K (3,3), links: indices into the URDF's link order (or
        names) -> mvp (Q, C, L, 4, 4) float32 on the device, computed by ehb_k_fk_mvp."""
        _dev_check(qpos, torch.float64, self.device, "qpos")
        Q, dof = qpos.shape
        cams = np.ascontiguousarray(np.asarray(cam_poses.cpu() if isinstance(cam_poses, torch.Tensor) else cam_poses, np.float64))
        if cams.ndim == 2:
            cams = cams[None]
        Kh = np.ascontiguousarray(np.asarray(K.cpu() if isinstance(K, torch.Tensor) else K, np.float32))
        sel = np.ascontiguousarray([robot["tree_index"][robot["link_names"][l] if not isinstance(l, str) else l] for l in links], np.int32)
        mvp = torch.empty((Q, len(cams), len(sel), 4, 4), dtype=torch.float32, device=self.device)
        _check(lib().ehb_explore_fk_mvp(self._h, robot["id"], _ptr(qpos), dof, Q, cams.ctypes.data_as(C.c_void_p), len(cams),
                                        Kh.ctypes.data_as(C.c_void_p), H, W, sel.ctypes.data_as(C.c_void_p), len(sel), _ptr(mvp),
                                        _stream(self.device)))
        return mvp

    def solver_step_host(self, mesh_ids, mvp_host, ref_dev, H, W, loss_host, g_mvp_host):
        """Host-buffer form: mvp_host (B,L,4,4) f32 pinned CPU tensor; loss_host (B,), g_mvp_host (B,L,4,4) f64
        pinned CPU tensors (written, stream synchronised on return)."""
        B, L = mvp_host.shape[0], mvp_host.shape[1]
        ids = (C.c_int * L)(*mesh_ids)
        _dev_check(ref_dev, torch.float32, self.device, "ref")
        _check(lib().ehb_solver_step_host(self._h, ids, L, B, _ptr(mvp_host), _ptr(ref_dev), H, W, _ptr(loss_host),
                                          _ptr(g_mvp_host), _stream(self.device)))

    def solver_step_host_u8(self, mesh_ids, mvp_host, ref_u8_host, H, W, loss_host, g_mvp_host):
        """Everything from host memory: mvp_host (B,L,4,4) f32 and ref_u8_host (B,H,W) u8 pinned CPU tensors are
        copied in, loss_host (B,) / g_mvp_host (B,L,4,4) f64 pinned CPU tensors are written; synchronises."""
        B, L = mvp_host.shape[0], mvp_host.shape[1]
        ids = (C.c_int * L)(*mesh_ids)
        if ref_u8_host.dtype != torch.uint8 or not ref_u8_host.is_contiguous():
            raise EhbError("ref_u8_host must be a contiguous uint8 CPU tensor")
        _check(lib().ehb_solver_step_host_u8(self._h, ids, L, B, _ptr(mvp_host), _ptr(ref_u8_host), H, W,
                                             _ptr(loss_host), _ptr(g_mvp_host), _stream(self.device)))

    # -- pose chain ----------------------------------------------------------------------------------------
    def pose_compose(self, dof, K, link_poses, H, W, out=None):
        """dof (6,), K (3,3), link_poses (B,L,4,4) f32 -> mvp (B,L,4,4) f32"""
        for t, n in ((dof, "dof"), (K, "K"), (link_poses, "link_poses")):
            _dev_check(t, torch.float32, self.device, n)
        B, L = link_poses.shape[0], link_poses.shape[1]
        if out is None:
            out = torch.empty((B, L, 4, 4), dtype=torch.float32, device=self.device)
        _check(lib().ehb_pose_compose(self._h, _ptr(dof), _ptr(K), _ptr(link_poses), B, L, H, W, _ptr(out),
                                      _stream(self.device)))
        return out

    def pose_backward(self, dof, K, link_poses, g_mvp, loss, H, W, grad_scale=1.0, loss_scale=None, out=None, send=False):
        """-> out7 f32 (7,) = [grad_scale * dL/ddof, loss_scale * sum(loss)].  send=True: the same kernel also posts out7 to
        every peer's mailbox (first half of the fused all-reduce; pair with adam_step(recv=True))."""
        _dev_check(g_mvp, torch.float64, self.device, "g_mvp")
        _dev_check(loss, torch.float64, self.device, "loss")
        B, L = link_poses.shape[0], link_poses.shape[1]
        if loss_scale is None:
            loss_scale = 1.0 / B
        if out is None:
            out = torch.empty((7,), dtype=torch.float32, device=self.device)
        fn = lib().ehb_pose_backward_send if send else lib().ehb_pose_backward
        _check(fn(self._h, _ptr(dof), _ptr(K), _ptr(link_poses), _ptr(g_mvp), _ptr(loss), B, L, H,
                  W, float(grad_scale), float(loss_scale), _ptr(out), _stream(self.device)))
        return out

    def adam_step(self, dof, g7, state, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, hist=None, recv=False, compose=None):
        """torch.optim.Adam step of the 6 parameters on the device.  compose=(K, link_poses, H, W, mvp_out): the same launch
        also writes the matrices of the next iteration from the updated parameters (ehb_adam_step_compose)."""
        _dev_check(dof, torch.float32, self.device, "dof")
        _dev_check(g7, torch.float32, self.device, "g7")
        _dev_check(state, torch.float32, self.device, "state")
        cap = 0 if hist is None else hist.shape[0]
        if compose is not None:
            K, lp, H, W, mvp = compose
            _dev_check(K, torch.float32, self.device, "K"); _dev_check(lp, torch.float32, self.device, "link_poses")
            _dev_check(mvp, torch.float32, self.device, "mvp")
            _check(lib().ehb_adam_step_compose(self._h, _ptr(dof), _ptr(g7), _ptr(state), lr, betas[0], betas[1], eps, weight_decay,
                                               _ptr(hist), cap, int(bool(recv)), _ptr(K), _ptr(lp), lp.shape[0], lp.shape[1], H, W,
                                               _ptr(mvp), _stream(self.device)))
            return
        fn = lib().ehb_adam_step_recv if recv else lib().ehb_adam_step
        _check(fn(self._h, _ptr(dof), _ptr(g7), _ptr(state), lr, betas[0], betas[1], eps, weight_decay,
                  _ptr(hist), cap, _stream(self.device)))

    def pose_backward_adam(self, dof, K, link_poses, g_mvp, loss, H, W, state, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                           grad_scale=1.0, loss_scale=None, out=None, exchange=False, adam_dof=None, hist=None, mvp_next=None):
        """pose_backward + (exchange) + adam_step (+ the next iteration's matrices) in one launch: the tail of a solver
        iteration.  adam_dof: the parameters Adam updates (default: dof itself)."""
        for t, n, dt in ((dof, "dof", torch.float32), (K, "K", torch.float32), (link_poses, "link_poses", torch.float32),
                         (g_mvp, "g_mvp", torch.float64), (loss, "loss", torch.float64), (state, "state", torch.float32)):
            _dev_check(t, dt, self.device, n)
        B, L = link_poses.shape[0], link_poses.shape[1]
        if loss_scale is None:
            loss_scale = 1.0 / B
        if out is None:
            out = torch.empty((7,), dtype=torch.float32, device=self.device)
        if adam_dof is None:
            adam_dof = dof
        _dev_check(adam_dof, torch.float32, self.device, "adam_dof")
        if mvp_next is not None:
            _dev_check(mvp_next, torch.float32, self.device, "mvp_next")
        cap = 0 if hist is None else hist.shape[0]
        _check(lib().ehb_pose_backward_adam(self._h, _ptr(dof), _ptr(K), _ptr(link_poses), _ptr(g_mvp), _ptr(loss), B, L, H, W,
                                            float(grad_scale), float(loss_scale), _ptr(out), int(bool(exchange)), _ptr(adam_dof),
                                            _ptr(state), lr, betas[0], betas[1], eps, weight_decay, _ptr(hist), cap,
                                            _ptr(mvp_next), _stream(self.device)))
        return out

    def solver_step_begin_u8(self, slot, mesh_ids, mvp_host, ref_u8_host, H, W, loss_host, g_mvp_host):
        """Asynchronous host-buffer step on slot 0/1 (own stream): returns at once; pair with solver_step_end(slot)."""
        B, L = mvp_host.shape[0], mvp_host.shape[1]
        ids = (C.c_int * L)(*mesh_ids)
        for t in (mvp_host, ref_u8_host, loss_host, g_mvp_host):
            if not t.is_pinned() or not t.is_contiguous():
                raise EhbError("host buffers of an asynchronous step must be pinned and contiguous")
        _check(lib().ehb_solver_step_begin_u8(self._h, slot, ids, L, B, _ptr(mvp_host), _ptr(ref_u8_host), H, W,
                                              _ptr(loss_host), _ptr(g_mvp_host)))

    def solver_step_begin_ref(self, slot, mesh_ids, mvp_host, ref: RefMasks, H, W, loss_host, g_mvp_host):
        """Asynchronous host-buffer step against registered reference masks: only the matrices go up."""
        B, L = mvp_host.shape[0], mvp_host.shape[1]
        ids = (C.c_int * L)(*mesh_ids)
        for t in (mvp_host, loss_host, g_mvp_host):
            if not t.is_pinned() or not t.is_contiguous():
                raise EhbError("host buffers of an asynchronous step must be pinned and contiguous")
        if len(ref) != B:
            raise EhbError("registered reference has %d views, the step %d" % (len(ref), B))
        _check(lib().ehb_solver_step_begin_ref(self._h, slot, ids, L, B, _ptr(mvp_host), ref.ref_id, ref.first, H, W,
                                               _ptr(loss_host), _ptr(g_mvp_host)))

    def step_begin(self, slot, mesh_ids, ref: RefMasks, H, W, mvp, masks=None, loss_host=None, g_mvp_host=None, dof=None,
                   K=None, link_poses=None, out7=None, out7_host=None, adam_dof=None, adam_state=None, lr=3e-3, weight_decay=0.0,
                   grad_scale=0.0, loss_scale=0.0, exchange=False):
        """General asynchronous step on slot 0..3 (ehb_step_begin): mvp (B,L,4,4) f32 is a pinned host tensor (copied in) or
        a device tensor; optional outputs masks (device), loss_host / g_mvp_host (pinned), and with dof / K / link_poses
        (device) the pose chain's out7 on the device (out7) and / or the host (out7_host, pinned); with adam_dof / adam_state
        the Adam update behind it, with exchange=True after the all-reduce of out7 over the connected ranks."""
        B, L = mvp.shape[0], mvp.shape[1]
        ids = (C.c_int * L)(*mesh_ids)
        if len(ref) != B or (ref.H, ref.W) != (H, W):
            raise EhbError("registered reference does not match the step (%d views of %dx%d)" % (len(ref), ref.H, ref.W))
        for t in (mvp if not mvp.is_cuda else None, loss_host, g_mvp_host, out7_host):
            if t is not None and (not t.is_pinned() or not t.is_contiguous()):
                raise EhbError("host buffers of an asynchronous step must be pinned and contiguous")
        io = StepIO()
        io.mvp_host, io.mvp_dev = (None, mvp.data_ptr()) if mvp.is_cuda else (mvp.data_ptr(), None)
        for name, t in (("masks_dev", masks), ("loss_host", loss_host), ("g_mvp_host", g_mvp_host), ("dof_dev", dof), ("K_dev", K),
                        ("link_poses_dev", link_poses), ("out7_dev", out7), ("out7_host", out7_host), ("adam_dof_dev", adam_dof),
                        ("adam_state_dev", adam_state)):
            setattr(io, name, None if t is None else t.data_ptr())
        io.grad_scale, io.loss_scale, io.lr, io.weight_decay, io.exchange = grad_scale, loss_scale, lr, weight_decay, int(bool(exchange))
        _check(lib().ehb_step_begin(self._h, slot, ids, L, B, ref.ref_id, ref.first, H, W, C.byref(io)))

    def slot_stream(self, slot):
        """The slot's CUDA stream as a torch.cuda.ExternalStream."""
        h = C.c_void_p()
        _check(lib().ehb_slot_stream(self._h, slot, C.byref(h)))
        return torch.cuda.ExternalStream(h.value, device=self.device)

    def slots_fork(self):
        """Every slot stream waits for the work enqueued so far on the current stream."""
        _check(lib().ehb_slots_fork(self._h, _stream(self.device)))

    def slots_join(self):
        """The current stream waits for everything enqueued on the slot streams."""
        _check(lib().ehb_slots_join(self._h, _stream(self.device)))

    def solver_step_end(self, slot):
        _check(lib().ehb_solver_step_end(self._h, slot))

    # -- NVLink one-shot all-reduce ----------------------------------------------------------------------------
    def comm_connect(self, group=None):
        """Exchange the CUDA IPC handles of the per-rank mailboxes through torch.distributed and map the peers."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        buf = (C.c_ubyte * 64)()
        _check(lib().ehb_comm_local_handle(self._h, buf))
        mine = torch.tensor(list(buf), dtype=torch.uint8, device=self.device)
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine, group=group)
        flat = torch.cat(allh).cpu().numpy().tobytes()
        _check(lib().ehb_comm_connect(self._h, rank, world, flat))
        dist.barrier(group)
        return rank, world

    def allreduce7(self, g7):
        _dev_check(g7, torch.float32, self.device, "g7")
        _check(lib().ehb_allreduce7(self._h, _ptr(g7), _stream(self.device)))
