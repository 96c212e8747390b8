"""numpy front end of the CPU SoftRas oracle (TEST INFRASTRUCTURE, not product code).

Mirrors the host side of the reference op:
  jrender/renderer/dr/softras/soft_rasterize.py:25     dist_eps -> ln(1/dist_eps - 1)
  jrender/renderer/dr/softras/soft_rasterize.py:39-42  enum maps
  jrender/renderer/dr/softras/soft_rasterize.py:59-71  buffer shapes
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
import ctypes as C

import numpy as np

from . import build as _build

DIST = {"hard": 0, "barycentric": 1, "euclidean": 2}
RGB = {"hard": 0, "softmax": 1, "none": 2}
ALPHA = {"hard": 0, "sum": 1, "prod": 2}
TEX = {"surface": 0, "vertex": 1}

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_build.build())
        f32p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_int32)
        _lib.softras_oracle_forward.restype = None
        _lib.softras_oracle_forward.argtypes = [
            f32p, f32p, f32p, f32p, f32p, i32p,
            C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
            C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float,
            C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        _lib.softras_oracle_backward.restype = None
        _lib.softras_oracle_backward.argtypes = [
            f32p, f32p, f32p, f32p, f32p, i32p, f32p, f32p, f32p,
            C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
            C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float,
            C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        _lib.softras_oracle_backward_mt.restype = None
        _lib.softras_oracle_backward_mt.argtypes = [
            f32p, f32p, f32p, f32p, f32p, i32p, f32p, f32p, f32p,
            C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
            C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float,
            C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        _lib.softras_oracle_max_threads.restype = C.c_int
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def dist_eps_logit(dist_eps):
    """soft_rasterize.py:25 -- float64 log, later narrowed to the kernel's float arg."""
    return np.float32(np.log(1.0 / dist_eps - 1.0))


class Params(dict):
    """Scalar parameters of the op with the reference's defaults (soft_rasterize.py:10-16)."""

    def __init__(self, **kw):
        d = dict(image_size=256, near=1.0, far=100.0, fill_back=True, eps=1e-3, sigma_val=1e-5,
                 dist_func="euclidean", dist_eps=1e-4, gamma_val=1e-4, aggr_func_rgb="softmax",
                 aggr_func_alpha="prod", texture_type="surface", max_faces_per_pixel_for_grad=16)
        unknown = set(kw) - set(d)
        if unknown:
            raise TypeError("unknown params %s" % sorted(unknown))
        d.update(kw)
        super().__init__(d)

    def scalars(self):
        return (np.float32(self["near"]), np.float32(self["far"]), np.float32(self["eps"]),
                np.float32(self["sigma_val"]), DIST[self["dist_func"]], dist_eps_logit(self["dist_eps"]),
                np.float32(self["gamma_val"]), RGB[self["aggr_func_rgb"]], ALPHA[self["aggr_func_alpha"]],
                TEX[self["texture_type"]], int(bool(self["fill_back"])))


def forward(face_vertices, textures, params, rows=None, nthreads=0, row_stride=1):
    """face_vertices [B,nf,3,3] f32, textures [B,nf,T,3] f32 ->
    dict(soft_colors [B,4,H,W], faces_info [B,nf,27], aggrs_info [B,2,H,W], faces_id_buffer [B,K,H,W] i32)."""
    fv = np.ascontiguousarray(face_vertices, dtype=np.float32)
    tx = np.ascontiguousarray(textures, dtype=np.float32)
    B, nf = fv.shape[:2]
    T = tx.shape[2]
    H = int(params["image_size"])
    K = int(params["max_faces_per_pixel_for_grad"])
    assert K <= 64, "reference hard limit kMaxPointsPerPixel (cuda/soft_rasterize.py:16)"
    faces_info = np.zeros((B, nf, 27), np.float32)
    aggrs_info = np.zeros((B, 2, H, H), np.float32)
    soft_colors = np.zeros((B, 4, H, H), np.float32)
    ids = np.zeros((B, K, H, H), np.int32)
    r0, r1 = (0, H) if rows is None else rows
    lib().softras_oracle_forward(
        _p(fv, C.c_float), _p(tx, C.c_float), _p(faces_info, C.c_float), _p(aggrs_info, C.c_float),
        _p(soft_colors, C.c_float), _p(ids, C.c_int32), B, nf, T, H, K, *params.scalars(), r0, r1, int(row_stride), nthreads)
    return dict(soft_colors=soft_colors, faces_info=faces_info, aggrs_info=aggrs_info, faces_id_buffer=ids)


def backward(face_vertices, textures, fwd, grad_soft_colors, params, accumulate_double=True, rows=None, row_stride=1,
             nthreads=1):
    """Top-K backward (K6).  `fwd` is the dict returned by forward().  -> (grad_faces, grad_textures).
    nthreads > 1 (CPU-baseline timing only): rows are dealt to threads with private double accumulators."""
    fv = np.ascontiguousarray(face_vertices, dtype=np.float32)
    tx = np.ascontiguousarray(textures, dtype=np.float32)
    g = np.ascontiguousarray(grad_soft_colors, dtype=np.float32)
    B, nf = fv.shape[:2]
    T = tx.shape[2]
    H = int(params["image_size"])
    K = int(params["max_faces_per_pixel_for_grad"])
    gf = np.zeros((B, nf, 3, 3), np.float32)
    gt = np.zeros((B, nf, T, 3), np.float32)
    r0, r1 = (0, H) if rows is None else rows
    if nthreads > 1:
        lib().softras_oracle_backward_mt(
            _p(fv, C.c_float), _p(tx, C.c_float), _p(np.ascontiguousarray(fwd["soft_colors"]), C.c_float),
            _p(np.ascontiguousarray(fwd["faces_info"]), C.c_float), _p(np.ascontiguousarray(fwd["aggrs_info"]), C.c_float),
            _p(np.ascontiguousarray(fwd["faces_id_buffer"]), C.c_int32), _p(g, C.c_float),
            _p(gf, C.c_float), _p(gt, C.c_float), B, nf, T, H, K, *params.scalars(), r0, r1, int(row_stride), int(nthreads))
        return gf, gt
    lib().softras_oracle_backward(
        _p(fv, C.c_float), _p(tx, C.c_float), _p(np.ascontiguousarray(fwd["soft_colors"]), C.c_float),
        _p(np.ascontiguousarray(fwd["faces_info"]), C.c_float), _p(np.ascontiguousarray(fwd["aggrs_info"]), C.c_float),
        _p(np.ascontiguousarray(fwd["faces_id_buffer"]), C.c_int32), _p(g, C.c_float),
        _p(gf, C.c_float), _p(gt, C.c_float), B, nf, T, H, K, *params.scalars(),
        int(bool(accumulate_double)), r0, r1, int(row_stride))
    return gf, gt


def max_threads():
    return int(lib().softras_oracle_max_threads())
