"""Drop-in for EasyHeC's ``NVDiffrastRenderer`` (easyhec/structures/nvdiffrast_renderer.py:10-73).

Same constructor and method signatures, same argument meaning, same return types:

    renderer = B200Renderer([H, W])
    mask = renderer.render_mask(verts, faces, K, object_pose, anti_aliasing=True)   # f32 (H,W) / bool (H,W)
    mask = renderer.batch_render_mask(verts, faces, K, anti_aliasing=True)

``dr.rasterize -> dr.interpolate(ones) -> dr.antialias -> [...,0] -> flip`` becomes one pass of
the sm_100a kernels in csrc/ through the C ABI; the backward re-rasterises instead of saving the rast /
colour / work-queue tensors (nothing but the final mask ever goes to HBM).  Gradients flow to
``object_pose`` / ``K`` through ``mvp`` and, when ``verts.requires_grad``, to the vertices.
"""
from collections import OrderedDict

import torch

from ._lib import Context, EhbError
from .projection import K_to_projection, opencv2gl

__all__ = ["B200Renderer", "NVDiffrastRenderer"]


class _RenderMaskAA(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mvp, verts, renderer, mesh_id):
        mvp_c = mvp.detach().contiguous().float()
        out = renderer.ctx.render_mask_fwd(mesh_id, mvp_c, renderer.H, renderer.W, True)
        ctx.renderer, ctx.mesh_id = renderer, mesh_id
        ctx.save_for_backward(mvp_c, verts)
        return out

    @staticmethod
    def backward(ctx, dy):
        mvp_c, verts = ctx.saved_tensors
        r = ctx.renderer
        want_v = ctx.needs_input_grad[1]
        # the registered mesh may have been re-posed by a later call: restore this call's vertices
        r._sync_verts(ctx.mesh_id, verts)
        g_mvp, g_pos = r.ctx.render_mask_bwd(ctx.mesh_id, mvp_c, r.H, r.W, dy.contiguous().float(), want_gpos=want_v)
        g_verts = (g_pos @ mvp_c[:, :3]).to(verts.dtype) if want_v else None
        return g_mvp.to(torch.float32), g_verts, None, None


class B200Renderer:
    def __init__(self, image_size, device=None):
        """image_size: H,W"""
        self.H, self.W = int(image_size[0]), int(image_size[1])
        self.resolution = image_size
        self.ctx = Context(device)
        self.device = self.ctx.device
        self.opencv2blender = opencv2gl(self.device)
        self._meshes = OrderedDict()   # (faces ptr, F, version) -> [mesh_id, faces ref, verts key, verts ref]
        self._max_cached = 64
        self._pf_key, self._pf = None, None   # projection @ flip of the last intrinsics tensor

    # -- mesh cache: the reference passes (verts, faces) tensors on every call ---------------------------
    @staticmethod
    def _key(t):
        return (t.data_ptr(), tuple(t.shape), t._version)

    def _check_inputs(self, verts, faces):
        if not (isinstance(verts, torch.Tensor) and verts.is_cuda and verts.dim() == 2 and verts.shape[1] == 3):
            raise EhbError("verts must be a CUDA float tensor of shape (N,3)")
        if not (isinstance(faces, torch.Tensor) and faces.is_cuda and faces.dim() == 2 and faces.shape[1] == 3):
            raise EhbError("faces must be a CUDA int32 tensor of shape (M,3)")
        if faces.dtype != torch.int32:
            raise EhbError("faces must have dtype int32, got %s" % faces.dtype)
        if verts.dtype != torch.float32:
            raise EhbError("verts must have dtype float32, got %s" % verts.dtype)

    def _mesh_for(self, verts, faces):
        self._check_inputs(verts, faces)
        faces = faces.contiguous()
        fk = self._key(faces) + (verts.shape[0],)
        ent = self._meshes.get(fk)
        vdet = verts.detach().contiguous()
        if ent is None:
            mid = self.ctx.register_mesh(vdet, faces)
            ent = [mid, faces, self._key(vdet), vdet]
            self._meshes[fk] = ent
            if len(self._meshes) > self._max_cached:
                _, old = self._meshes.popitem(last=False)
                self.ctx.release_mesh(old[0])
        else:
            self._meshes.move_to_end(fk)
            self._sync_verts(ent[0], vdet, ent)
        return ent[0]

    def _sync_verts(self, mesh_id, verts, ent=None):
        if ent is None:
            ent = next((e for e in self._meshes.values() if e[0] == mesh_id), None)
            if ent is None:
                return
        vdet = verts.detach().contiguous()
        k = self._key(vdet)
        if k != ent[2]:
            self.ctx.update_verts(mesh_id, vdet)
            ent[2], ent[3] = k, vdet

    # -- the operator ------------------------------------------------------------------------------------
    def _render(self, verts, faces, mvp, anti_aliasing):
        self.ctx.check("render_mask")        # an earlier launch overflowed its scratch: raise instead of returning garbage
        mesh_id = self._mesh_for(verts, faces)
        if anti_aliasing:
            return _RenderMaskAA.apply(mvp, verts, self, mesh_id)
        out = self.ctx.render_mask_fwd(mesh_id, mvp.detach().contiguous().float(), self.H, self.W, False)
        return out.view(torch.bool)

    def render_mask(self, verts, faces, K, object_pose, anti_aliasing=True):
        """
        @param verts: N,3, torch.tensor, float, cuda
        @param faces: M,3, torch.tensor, int32, cuda
        @param K: 3,3 torch.tensor, float ,cuda
        @param object_pose: 4,4 torch.tensor, float, cuda
        @return: mask: 0 to 1, HxW torch.cuda.FloatTensor (bool when anti_aliasing=False)
        """
        return self._render(verts, faces, self._proj_flip(K) @ object_pose, anti_aliasing)

    def batch_render_mask(self, verts, faces, K, anti_aliasing=True):
        """Vertices already in the camera frame (packed multi-link mesh, render_api.py:81-92)."""
        return self._render(verts, faces, self._proj_flip(K), anti_aliasing)

    def _proj_flip(self, K):
        """``K_to_projection(K, H, W) @ diag(1,-1,-1,1)``, kept per intrinsics tensor (storage + version): the reference
        rebuilds the projection from sixteen 0-d tensors on every call (nvdiffrast_renderer.py:33-36), ~20 micro-kernels per
        (view, link).  Multiplying by the flip first is bit-identical to the reference's ``proj @ (flip @ pose)``: the flip
        only changes signs."""
        if isinstance(K, torch.Tensor) and K.requires_grad:
            return K_to_projection(K, self.H, self.W).to(self.device) @ self.opencv2blender
        key = (K.data_ptr(), K._version, str(K.device)) if isinstance(K, torch.Tensor) else None
        if key is None or key != self._pf_key:
            self._pf = K_to_projection(K, self.H, self.W).to(self.device) @ self.opencv2blender
            self._pf_key = key
        return self._pf


NVDiffrastRenderer = B200Renderer
