"""CPU oracle of the SoftRas texture bake (TEST INFRASTRUCTURE; product code never imports it).

numpy restatement of load_textures_cuda_kernel, jrender/io/utils/load_textures.py:11-69 (barycentric sample point per
texel :29-37, UV -> image position :42-45, bilinear fetch :46-58, same float / double promotions).  Pinned against the
reference's own kernel compiled from /root/reference (oracle/build_ref.py) and run on a B200:
tests/golden/ref_gpu_bake_random40_R5.npz (generator oracle/make_ref_golden.py --bake-only).
"""
import numpy as np


def bake_textures_for_softras(image, faces_uv, textures, is_update):
    """load_textures.py:11-69.

    image [H,W,3] (already flipped vertically by the caller), faces_uv [nf,3,2],
    textures [nf,R*R,3] (returned updated where is_update != 0)."""
    nf, T = textures.shape[:2]
    R = int(np.sqrt(T))
    H, W = image.shape[:2]
    wy, wx = np.divmod(np.arange(T), R)
    lower = (wx + wy) < R
    w0 = np.where(lower, (wx + 1. / 3.) / R, ((R - 1. - wx) + 2. / 3.) / R).astype(np.float32)
    w1 = np.where(lower, (wy + 1. / 3.) / R, ((R - 1. - wy) + 2. / 3.) / R).astype(np.float32)
    w2 = (1. - w0.astype(np.float64) - w1.astype(np.float64)).astype(np.float32)
    f = faces_uv.astype(np.float32)
    pos_x = ((f[:, None, 0, 0] * w0 + f[:, None, 1, 0] * w1 + f[:, None, 2, 0] * w2) * np.float32(W - 1)).astype(np.float32)
    pos_y = ((f[:, None, 0, 1] * w0 + f[:, None, 1, 1] * w1 + f[:, None, 2, 1] * w2) * np.float32(H - 1)).astype(np.float32)
    ix, iy = pos_x.astype(np.int64), pos_y.astype(np.int64)       # C truncation; UVs are >= 0
    wx1 = pos_x - ix
    wx0 = 1 - wx1
    wy1 = pos_y - iy
    wy0 = 1 - wy1
    flat = image.reshape(-1, 3)
    n = flat.shape[0]
    iy1 = (pos_y + 1).astype(np.int64)

    def px(yy, xx):  # the reference indexes the flat buffer without clamping; keep in-bounds here
        return flat[np.clip(yy * W + xx, 0, n - 1)]
    c = (px(iy, ix) * (wx0 * wy0)[..., None] + px(iy1, ix) * (wx0 * wy1)[..., None] +
         px(iy, ix + 1) * (wx1 * wy0)[..., None] + px(iy1, ix + 1) * (wx1 * wy1)[..., None]).astype(np.float32)
    out = textures.copy()
    m = np.asarray(is_update) != 0
    out[m] = c[m]
    return out


