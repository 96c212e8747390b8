"""numpy front end of the CPU NMR (n3mr) oracle (TEST INFRASTRUCTURE, not product code).

Mirrors the host side of the reference Function, jrender/renderer/dr/n3mr/n3mr.py:
  :69-123  execute (buffer shapes, sub-op order, background mix :135-143, alpha :145-148)
  :29-67   grad    (backward_pixel_map -> backward_textures -> backward_depth_map in place)
Maps keep the kernels' orientation: [B, yi, xi] with yi up (row 0 = bottom); the flip to image
orientation belongs to rasterize_rgbad (n3mr.py:239-247).
"""
import ctypes as C

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_build.build())
        f, i, I, F = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_int, C.c_float
        L.nmr_oracle_face_index_map.restype = None
        L.nmr_oracle_face_index_map.argtypes = [f, i, f, f, f, I, I, I, F, F, I]
        L.nmr_oracle_texture_sampling.restype = None
        L.nmr_oracle_texture_sampling.argtypes = [f, f, i, f, f, f, i, f, I, I, I, I, F]
        L.nmr_oracle_backward_pixel_map.restype = None
        L.nmr_oracle_backward_pixel_map.argtypes = [f, i, f, f, f, f, f, I, I, I, F, I, I]
        L.nmr_oracle_backward_textures.restype = None
        L.nmr_oracle_backward_textures.argtypes = [i, f, i, f, f, I, I, I, I]
        L.nmr_oracle_backward_depth_map.restype = None
        L.nmr_oracle_backward_depth_map.argtypes = [f, f, i, f, f, f, f, I, I, I]
        _lib = L
    return _lib


def _f(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def forward(faces, textures, image_size, near=0.1, far=100.0, eps=1e-4, background_color=(0, 0, 0),
            return_rgb=True, return_alpha=True, return_depth=True):
    """faces [B,nf,3,3], textures [B,nf,ts,ts,ts,3] or None -> dict of maps (kernel orientation)."""
    L = lib()
    fc = np.ascontiguousarray(faces, np.float32)
    B, nf = fc.shape[:2]
    H = int(image_size)
    out = dict(face_index_map=np.empty((B, H, H), np.int32), weight_map=np.empty((B, H, H, 3), np.float32),
               depth_map=np.empty((B, H, H), np.float32),
               face_inv_map=np.zeros((B, H, H, 3, 3), np.float32) if return_depth else None)
    L.nmr_oracle_face_index_map(_f(fc), _i(out["face_index_map"]), _f(out["weight_map"]), _f(out["depth_map"]),
                                _f(out["face_inv_map"]) if return_depth else None, B, nf, H,
                                np.float32(near), np.float32(far), int(return_depth))
    if return_rgb:
        tx = np.ascontiguousarray(textures, np.float32)
        ts = tx.shape[2]
        out["rgb_raw"] = np.empty((B, H, H, 3), np.float32)
        out["sampling_index_map"] = np.empty((B, H, H, 8), np.int32)
        out["sampling_weight_map"] = np.empty((B, H, H, 8), np.float32)
        L.nmr_oracle_texture_sampling(_f(fc), _f(tx), _i(out["face_index_map"]), _f(out["weight_map"]),
                                      _f(out["depth_map"]), _f(out["rgb_raw"]), _i(out["sampling_index_map"]),
                                      _f(out["sampling_weight_map"]), B, nf, H, ts, np.float32(eps))
        mask = (out["face_index_map"] >= 0).astype(np.float32)[..., None]
        bg = np.asarray(background_color, np.float32)
        bg = bg[None, None, None, :] if bg.ndim == 1 else bg[:, None, None, :]
        out["rgb_map"] = (out["rgb_raw"] * mask + (1 - mask) * bg).astype(np.float32)     # n3mr.py:135-143
    if return_alpha:
        out["alpha_map"] = (out["face_index_map"] >= 0).astype(np.float32)                 # :145-148
    return out


def backward(faces, textures, fwd, image_size, eps=1e-4, grad_rgb_map=None, grad_alpha_map=None,
             grad_depth_map=None, return_rgb=True, return_alpha=True, return_depth=True):
    """-> (grad_faces [B,nf,3,3], grad_textures or None), following n3mr.py:29-67."""
    L = lib()
    fc = np.ascontiguousarray(faces, np.float32)
    B, nf = fc.shape[:2]
    H = int(image_size)
    grad_faces = np.zeros((B, nf, 3, 3), np.float32)
    one = np.zeros(1, np.float32)
    g_rgb = np.ascontiguousarray(grad_rgb_map, np.float32) if return_rgb else one
    g_a = np.ascontiguousarray(grad_alpha_map, np.float32) if return_alpha else one
    if return_rgb or return_alpha:
        L.nmr_oracle_backward_pixel_map(_f(fc), _i(fwd["face_index_map"]),
                                        _f(fwd["rgb_map"]) if return_rgb else _f(one),
                                        _f(fwd["alpha_map"]) if return_alpha else _f(one), _f(g_rgb), _f(g_a),
                                        _f(grad_faces), B, nf, H, np.float32(eps), int(return_rgb), int(return_alpha))
    grad_textures = None
    if return_rgb:
        tx = np.ascontiguousarray(textures, np.float32)
        grad_textures = np.zeros_like(tx)
        L.nmr_oracle_backward_textures(_i(fwd["face_index_map"]), _f(fwd["sampling_weight_map"]),
                                       _i(fwd["sampling_index_map"]), _f(g_rgb), _f(grad_textures), B, nf, H, tx.shape[2])
    if return_depth:
        g_d = np.ascontiguousarray(grad_depth_map, np.float32)
        L.nmr_oracle_backward_depth_map(_f(fc), _f(fwd["depth_map"]), _i(fwd["face_index_map"]),
                                        _f(np.ascontiguousarray(fwd["face_inv_map"])), _f(fwd["weight_map"]),
                                        _f(g_d), _f(grad_faces), B, nf, H)
    return grad_faces, grad_textures
