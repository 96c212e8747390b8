"""Run the reference's OWN kernels (oracle/_ref/libjrender_ref.so, built by oracle/build_ref.py
from the sources under /root/reference) on the GPU.  TEST INFRASTRUCTURE, not product code.

Host semantics mirrored from jrender/renderer/dr/softras/soft_rasterize.py:34-133 (buffer
shapes, the [B,K,H,W] -> [B,H,W,K] transpose before the backward, max_elems_per_bin = nf/5).
"""
import ctypes as C
import os

import numpy as np

from . import build_ref
from .softras import Params  # noqa: F401  (same parameter object as the CPU oracle)

_lib = None


def available():
    return os.path.exists(build_ref.LIB)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libjrender_ref.so missing: run `python -m oracle.build_ref` where /root/reference exists")
        L = C.CDLL(build_ref.LIB)
        P, I, F = C.c_void_p, C.c_int, C.c_float
        scal = [I, I, I, I, I, F, F, F, F, I, F, F, I, I, I, I]
        L.ref_softras_forward.restype = I
        L.ref_softras_forward.argtypes = [P] * 6 + scal + [P]
        L.ref_softras_backward.restype = I
        L.ref_softras_backward.argtypes = [P] * 9 + scal + [P]
        L.ref_softras_forward_c2f.restype = I
        L.ref_softras_forward_c2f.argtypes = [P] * 6 + scal + [I, I]
        _lib = L
    return _lib


def _scalars(params, B, nf, T):
    near, far, eps, sigma, dist, dist_eps, gamma, rgb, alpha, tex, ds = params.scalars()
    return (B, nf, T, int(params["image_size"]), int(params["max_faces_per_pixel_for_grad"]),
            float(near), float(far), float(eps), float(sigma), dist, float(dist_eps), float(gamma), rgb, alpha, tex, ds)


def forward_t(fv, tex, params, c2f_bin_size=0, max_elems_per_bin=0):
    """torch CUDA tensors in -> dict of torch CUDA tensors (reference layout, ids [B,K,H,W])."""
    import torch
    L = lib()
    B, nf = fv.shape[:2]
    T = tex.shape[2]
    H, K = int(params["image_size"]), int(params["max_faces_per_pixel_for_grad"])
    dev = fv.device
    out = dict(faces_info=torch.empty((B, nf, 27), dtype=torch.float32, device=dev),
               aggrs_info=torch.empty((B, 2, H, H), dtype=torch.float32, device=dev),
               soft_colors=torch.empty((B, 4, H, H), dtype=torch.float32, device=dev),
               faces_id_buffer=torch.empty((B, K, H, H), dtype=torch.int32, device=dev))
    ptrs = [C.c_void_p(t.data_ptr()) for t in (fv, tex, out["faces_info"], out["aggrs_info"], out["soft_colors"], out["faces_id_buffer"])]
    if c2f_bin_size:
        m = max_elems_per_bin or int(nf / 5)          # soft_rasterize.py:85-87
        rc = L.ref_softras_forward_c2f(*ptrs, *_scalars(params, B, nf, T), int(c2f_bin_size), int(m))
    else:
        rc = L.ref_softras_forward(*ptrs, *_scalars(params, B, nf, T), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if rc != 0:
        raise RuntimeError("reference forward kernel launch failed: cudaError %d" % rc)
    return out


def backward_t(fv, tex, fwd, grad, params, ids_bhwk=None):
    import torch
    L = lib()
    B, nf = fv.shape[:2]
    T = tex.shape[2]
    dev = fv.device
    if ids_bhwk is None:
        ids_bhwk = fwd["faces_id_buffer"].permute(0, 2, 3, 1).contiguous()   # soft_rasterize.py:108
    gf = torch.empty((B, nf, 3, 3), dtype=torch.float32, device=dev)
    gt = torch.empty((B, nf, T, 3), dtype=torch.float32, device=dev)
    ptrs = [C.c_void_p(t.data_ptr()) for t in (fv, tex, fwd["soft_colors"], fwd["faces_info"], fwd["aggrs_info"], ids_bhwk, grad, gf, gt)]
    rc = L.ref_softras_backward(*ptrs, *_scalars(params, B, nf, T), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if rc != 0:
        raise RuntimeError("reference backward kernel launch failed: cudaError %d" % rc)
    return gf, gt


def run(fv, tex, params, grad=None, device="cuda:0"):
    """numpy in -> numpy out, same dict layout as oracle.softras / tests.util.run_cuda."""
    import torch
    dev = torch.device(device)
    fvt = torch.from_numpy(np.ascontiguousarray(fv, dtype=np.float32)).to(dev)
    txt = torch.from_numpy(np.ascontiguousarray(tex, dtype=np.float32)).to(dev)
    fwd = forward_t(fvt, txt, params)
    out = {k: v.cpu().numpy() for k, v in fwd.items()}
    if grad is not None:
        g = torch.from_numpy(np.ascontiguousarray(grad, dtype=np.float32)).to(dev)
        gf, gt = backward_t(fvt, txt, fwd, g, params)
        out["grad_faces"], out["grad_textures"] = gf.cpu().numpy(), gt.cpu().numpy()
    torch.cuda.synchronize()
    return out
