"""ctypes front end of oracle/ehb_oracle.c (CPU restatement of the reference's render_mask path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (easyhec_b200/) never imports it.
Parity status: host geometry pinned by reference-run goldens; rasterize/antialias semantics
restate un-vendored nvdiffrast => "parity unpinned" there (see ehb_oracle.c header).
"""
import ctypes as C
import os
import subprocess

import numpy as np

__all__ = ["build", "lib", "num_threads", "set_num_threads", "transform", "build_adjacency", "rasterize", "render_mask",
           "render_mask_bwd", "render_views", "union_binary", "variance_scores", "pack_links", "FILL_RULE"]

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libehb_oracle.so")
_lib = None
FILL_RULE = 0


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ehb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libehb_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.eho_num_threads.restype = C.c_int
        for f in ("eho_rasterize", "eho_render_mask", "eho_render_views", "eho_union_binary"):
            getattr(_lib, f).restype = C.c_int
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def num_threads() -> int:
    return int(lib().eho_num_threads())


def set_num_threads(n: int) -> int:
    """Threads of the view-parallel loops (OpenMP); returns the count now in effect."""
    lib().eho_set_num_threads(C.c_int(int(n)))
    return num_threads()


def transform(verts, mvp):
    verts, mvp = _f32(verts), _f32(mvp)
    out = np.empty((len(verts), 4), np.float32)
    lib().eho_transform(_p(verts, C.c_float), C.c_int(len(verts)), _p(mvp, C.c_float), _p(out, C.c_float))
    return out


def build_adjacency(faces, V):
    faces = _i32(faces)
    opp = np.empty((len(faces), 3), np.int32)
    lib().eho_build_adjacency(_p(faces, C.c_int), C.c_int(len(faces)), C.c_int(V), _p(opp, C.c_int))
    return opp


def rasterize(clip, faces, H, W, rule=None):
    """-> (tri_id (H,W) int64 GL rows, -1 empty; zkey (H,W) uint32; n_need_clip)"""
    clip, faces = _f32(clip), _i32(faces)
    key = np.empty((H, W), np.uint64)
    nclip = lib().eho_rasterize(_p(clip, C.c_float), C.c_int(len(clip)), _p(faces, C.c_int), C.c_int(len(faces)),
                                C.c_int(H), C.c_int(W), C.c_int(FILL_RULE if rule is None else rule),
                                _p(key, C.c_uint64))
    empty = key == np.uint64(0xFFFFFFFFFFFFFFFF)
    tid = (key & np.uint64(0xFFFFFFFF)).astype(np.int64)
    tid[empty] = -1
    return tid, (key >> np.uint64(32)).astype(np.uint32), nclip


def render_mask(verts, faces, mvp, H, W, anti_aliasing=True, opp=None, rule=None, save=False):
    """The single-mesh operator on CPU.  Returns mask (H,W) float32 (AA) or bool; with save=True also
    the state tuple the backward needs."""
    verts, faces, mvp = _f32(verts), _i32(faces), _f32(mvp)
    V, F = len(verts), len(faces)
    if opp is None:
        opp = build_adjacency(faces, V)
    opp = _i32(opp)
    n = H * W
    of = np.zeros((H, W), np.float32) if anti_aliasing else None
    ou = None if anti_aliasing else np.zeros((H, W), np.uint8)
    key = np.empty(n, np.uint64) if save else None
    alpha = np.empty(2 * n, np.float32) if save and anti_aliasing else None
    info = np.empty(2 * n, np.uint8) if save and anti_aliasing else None
    nclip = lib().eho_render_mask(_p(verts, C.c_float), C.c_int(V), _p(faces, C.c_int), C.c_int(F), _p(opp, C.c_int),
                                  _p(mvp, C.c_float), C.c_int(H), C.c_int(W), C.c_int(int(anti_aliasing)),
                                  C.c_int(FILL_RULE if rule is None else rule), _p(of, C.c_float), _p(ou, C.c_uint8),
                                  _p(key, C.c_uint64), _p(alpha, C.c_float), _p(info, C.c_uint8))
    out = of if anti_aliasing else ou.astype(bool)
    if save:
        return out, (key, alpha, info, nclip)
    return out


def render_mask_bwd(verts, faces, mvp, H, W, state, dy):
    """dy (H,W) image rows -> (g_pos (V,4) float64, g_mvp (4,4) float64)."""
    verts, faces, mvp, dy = _f32(verts), _i32(faces), _f32(mvp), _f32(dy)
    key, alpha, info, _ = state
    gpos = np.zeros((len(verts), 4), np.float64)
    gmvp = np.zeros((4, 4), np.float64)
    lib().eho_render_mask_bwd(_p(verts, C.c_float), C.c_int(len(verts)), _p(faces, C.c_int), C.c_int(len(faces)),
                              _p(mvp, C.c_float), C.c_int(H), C.c_int(W), _p(key, C.c_uint64), _p(alpha, C.c_float),
                              _p(info, C.c_uint8), _p(dy, C.c_float), _p(gpos, C.c_double), _p(gmvp, C.c_double))
    return gpos, gmvp


def pack_links(meshes):
    """[(verts, faces)] or Mesh objects -> dict(verts, voff, faces, foff, opp): concatenated arrays,
    faces local to each link."""
    vs, fs, ops, voff, foff = [], [], [], [0], [0]
    for m in meshes:
        v, f = (m.vertices, m.faces) if hasattr(m, "vertices") else m
        v, f = _f32(v), _i32(f)
        vs.append(v); fs.append(f); ops.append(build_adjacency(f, len(v)))
        voff.append(voff[-1] + len(v)); foff.append(foff[-1] + len(f))
    return dict(verts=np.concatenate(vs), faces=np.concatenate(fs), opp=np.concatenate(ops),
                voff=np.asarray(voff, np.int32), foff=np.asarray(foff, np.int32), L=len(vs))


def render_views(packed, mvp, ref, H, W, rule=None, backward=True):
    """RBSolver mask loop on CPU.  mvp (B,L,4,4), ref (B,H,W) float -> dict(masks (B,H,W) f32,
    loss_per_view (B,) f64, loss f64 (mean over views), g_mvp (B,L,4,4) f64, n_need_clip)."""
    mvp, ref = _f32(mvp), _f32(ref)
    B, L = mvp.shape[0], mvp.shape[1]
    assert L == packed["L"]
    masks = np.empty((B, H, W), np.float32)
    loss_b = np.zeros(B, np.float64)
    gmvp = np.zeros((B, L, 4, 4), np.float64)
    nclip = lib().eho_render_views(C.c_int(B), C.c_int(L), _p(packed["verts"], C.c_float), _p(packed["voff"], C.c_int),
                                   _p(packed["faces"], C.c_int), _p(packed["opp"], C.c_int), _p(packed["foff"], C.c_int),
                                   _p(mvp, C.c_float), _p(ref, C.c_float), C.c_int(H), C.c_int(W),
                                   C.c_int(FILL_RULE if rule is None else rule), C.c_int(int(backward)),
                                   _p(masks, C.c_float), _p(loss_b, C.c_double), _p(gmvp, C.c_double))
    return dict(masks=masks, loss_per_view=loss_b, loss=loss_b.mean() if B else 0.0, g_mvp=gmvp, n_need_clip=nclip)


def union_binary(packed, mvp, H, W, rule=None):
    """N renders of the packed robot, one depth buffer per render.  mvp (N,L,4,4) -> bool (N,H,W)."""
    mvp = _f32(mvp)
    N, L = mvp.shape[0], mvp.shape[1]
    out = np.zeros((N, H, W), np.uint8)
    lib().eho_union_binary(C.c_int(N), C.c_int(L), _p(packed["verts"], C.c_float), _p(packed["voff"], C.c_int),
                           _p(packed["faces"], C.c_int), _p(packed["foff"], C.c_int), _p(mvp, C.c_float),
                           C.c_int(H), C.c_int(W), C.c_int(FILL_RULE if rule is None else rule), _p(out, C.c_uint8))
    return out.astype(bool)


def variance_scores(masks):
    """masks (Q,C,H,W) bool -> (Q,) float64: sum over pixels of the unbiased variance over C."""
    m = np.ascontiguousarray(masks, dtype=np.uint8)
    Q, Cn = m.shape[0], m.shape[1]
    out = np.zeros(Q, np.float64)
    lib().eho_variance_scores(_p(m, C.c_uint8), C.c_int(Q), C.c_int(Cn), C.c_long(m.shape[2] * m.shape[3]),
                              _p(out, C.c_double))
    return out
