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


# --------------------------------------------------------------------------- NMR (dr_type='n3mr')
# Host semantics mirrored from jrender/renderer/dr/n3mr/n3mr.py:29-67 (grad), :69-123 (execute),
# :135-148 (background mix, alpha): buffer shapes, the dummy 1-element tensors of disabled outputs,
# the order pixel-map -> textures -> depth of the backward.
def _nmr_lib():
    L = lib()
    if not hasattr(L, "_nmr_ready"):
        P, I, F = C.c_void_p, C.c_int, C.c_float
        L.ref_nmr_forward_face_index_map.restype = I
        L.ref_nmr_forward_face_index_map.argtypes = [P] * 7 + [I, I, I, F, F, I, I, I, C.c_long]
        L.ref_nmr_forward_texture_sampling.restype = I
        L.ref_nmr_forward_texture_sampling.argtypes = [P] * 8 + [I, I, I, I, F]
        L.ref_nmr_backward_pixel_map.restype = I
        L.ref_nmr_backward_pixel_map.argtypes = [P] * 7 + [I, I, I, F, I, I]
        L.ref_nmr_backward_textures.restype = I
        L.ref_nmr_backward_textures.argtypes = [P] * 5 + [I, I, I, I]
        L.ref_nmr_backward_depth_map.restype = I
        L.ref_nmr_backward_depth_map.argtypes = [P] * 7 + [I, I, I]
        L._nmr_ready = True
    return L


def nmr_run(faces, textures, image_size, near=0.1, far=100.0, eps=1e-4, background_color=(0, 0, 0),
            flags=(True, True, True), grads=None, device="cuda:0"):
    """numpy in -> dict of numpy maps in the kernels' orientation (same keys as oracle.nmr.forward) and,
    with grads=(grad_rgb, grad_alpha, grad_depth), grad_faces / grad_textures."""
    import torch
    L = _nmr_lib()
    dev = torch.device(device)
    rrgb, ralpha, rdepth = (bool(f) for f in flags)
    fc = torch.from_numpy(np.ascontiguousarray(faces, np.float32)).to(dev)
    B, nf = fc.shape[:2]
    H = int(image_size)
    one_f = lambda: torch.zeros(1, dtype=torch.float32, device=dev)   # noqa: E731
    tx = torch.from_numpy(np.ascontiguousarray(textures, np.float32)).to(dev) if rrgb else one_f()
    ts = int(tx.shape[2]) if rrgb else 0
    p = lambda t: C.c_void_p(t.data_ptr())                              # noqa: E731
    fim = torch.empty((B, H, H), dtype=torch.int32, device=dev)
    wm = torch.empty((B, H, H, 3), dtype=torch.float32, device=dev)
    dm = torch.empty((B, H, H), dtype=torch.float32, device=dev)
    finv = torch.empty((B, H, H, 3, 3), dtype=torch.float32, device=dev) if rdepth else one_f()
    faces_inv = torch.zeros_like(fc)                                    # n3mr.py:126
    lock = torch.empty((B, H, H), dtype=torch.int32, device=dev)
    rc = L.ref_nmr_forward_face_index_map(p(fc), p(faces_inv), p(fim), p(wm), p(dm), p(finv), p(lock), B, nf, H,
                                          float(near), float(far), int(rrgb), int(ralpha), int(rdepth), finv.numel())
    if rc != 0:
        raise RuntimeError("reference forward_face_index_map failed: %d (image sizes instantiated: 32, 48, 64, 256, 512, 1024)" % rc)
    out = dict(face_index_map=fim, weight_map=wm, depth_map=dm)
    if rdepth:
        out["face_inv_map"] = finv
    if rrgb:
        rgb = torch.empty((B, H, H, 3), dtype=torch.float32, device=dev)
        sim = torch.empty((B, H, H, 8), dtype=torch.int32, device=dev)
        swm = torch.empty((B, H, H, 8), dtype=torch.float32, device=dev)
        rc = L.ref_nmr_forward_texture_sampling(p(fc), p(tx), p(fim), p(wm), p(dm), p(rgb), p(sim), p(swm), B, nf, H, ts, float(eps))
        if rc != 0:
            raise RuntimeError("reference forward_texture_sampling failed: %d" % rc)
        mask = (fim >= 0).float().unsqueeze(-1)
        bg = torch.tensor(np.asarray(background_color, np.float32), device=dev)
        bg = bg[None, None, None, :] if bg.dim() == 1 else bg[:, None, None, :]
        out["rgb_raw"] = rgb
        out["rgb_map"] = rgb * mask + (1 - mask) * bg                   # n3mr.py:135-143
        out["sampling_index_map"], out["sampling_weight_map"] = sim, swm
    if ralpha:
        out["alpha_map"] = (fim >= 0).float()                           # :145-148
    res = {k: v.cpu().numpy() for k, v in out.items()}
    if grads is not None:
        g_rgb, g_a, g_d = grads
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.float32)).to(dev)   # noqa: E731
        g_rgb = t(g_rgb) if rrgb else one_f()
        g_a = t(g_a) if ralpha else one_f()
        g_d = t(g_d) if rdepth else one_f()
        gf = torch.empty((B, nf, 3, 3), dtype=torch.float32, device=dev)
        rgb_map = out["rgb_map"].contiguous() if rrgb else one_f()
        alpha_map = out["alpha_map"].contiguous() if ralpha else one_f()
        rc = L.ref_nmr_backward_pixel_map(p(fc), p(fim), p(rgb_map), p(alpha_map), p(g_rgb), p(g_a), p(gf), B, nf, H,
                                          float(eps), int(rrgb), int(ralpha))
        if rc != 0:
            raise RuntimeError("reference backward_pixel_map failed: %d" % rc)
        if rrgb:
            gt = torch.empty_like(tx)
            rc = L.ref_nmr_backward_textures(p(fim), p(out["sampling_weight_map"]), p(out["sampling_index_map"]), p(g_rgb), p(gt),
                                             B, nf, H, ts)
            if rc != 0:
                raise RuntimeError("reference backward_textures failed: %d" % rc)
            res["grad_textures"] = gt.cpu().numpy()
        if rdepth:   # the kernel returns immediately per pixel without a depth gradient; called as the reference does
            rc = L.ref_nmr_backward_depth_map(p(fc), p(dm), p(fim), p(finv), p(wm), p(g_d), p(gf), B, nf, H)
            if rc != 0:
                raise RuntimeError("reference backward_depth_map failed: %d" % rc)
        res["grad_faces"] = gf.cpu().numpy()
    torch.cuda.synchronize()
    return res
