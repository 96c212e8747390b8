"""CPU oracle of the mesh regularisers (TEST INFRASTRUCTURE; product code never imports it).

float64 numpy restatement, expression by expression, of
    jrender/loss/laplacian_loss.py:5-36   LaplacianLoss (dense nv x nv matrix exactly as :11-26 builds it)
    jrender/loss/flatten_loss.py:5-79     FlattenLoss   (edge table exactly as :12-30 builds it, O(E*F) scan included)
and their gradients by central differences in float64 -- no autograd, no hand-derived formula, so the check of the fused
CUDA kernels' gradients (csrc/mesh_loss_api.cu) is independent of the PyTorch mirror in jrender_b200/loss.py.
Pinned: tests/test_host_golden.py holds `laplacian()` / `flatten()` to the values the reference's own Python produced
(tests/golden/ref_host_loss_sphere280.npz, generator oracle/make_ref_host_golden.py).
"""
import numpy as np


def laplacian_matrix(nv, faces):
    """laplacian_loss.py:11-26."""
    lap = np.zeros([nv, nv], np.float32)
    f = np.asarray(faces).astype(np.int64)
    lap[f[:, 0], f[:, 1]] = -1
    lap[f[:, 1], f[:, 0]] = -1
    lap[f[:, 1], f[:, 2]] = -1
    lap[f[:, 2], f[:, 1]] = -1
    lap[f[:, 2], f[:, 0]] = -1
    lap[f[:, 0], f[:, 2]] = -1
    r, c = np.diag_indices(nv)
    lap[r, c] = -lap.sum(1)
    for i in range(nv):
        lap[i, :] /= lap[i, i]
    return lap.astype(np.float64)


def laplacian(x, lap):
    """laplacian_loss.py:30-34: x [B,nv,3] -> [B]."""
    y = np.einsum("ij,bjk->bik", lap, np.asarray(x, np.float64))
    return (y ** 2).sum((1, 2))


def flatten_edges(faces):
    """flatten_loss.py:11-30 -> v0s, v1s, v2s, v3s (the reference's own scan; set order replaced by sorted order,
    which only permutes the terms of a sum)."""
    f = np.asarray(faces).astype(np.int64)
    pairs = sorted(set(tuple(v) for v in np.sort(np.concatenate((f[:, 0:2], f[:, 1:3]), axis=0))))
    v0s = np.array([p[0] for p in pairs], np.int64)
    v1s = np.array([p[1] for p in pairs], np.int64)
    v2s, v3s = [], []
    for v0, v1 in zip(v0s, v1s):
        count = 0
        for face in f:
            if v0 in face and v1 in face:
                v = face[(face != v0) & (face != v1)]
                if count == 0:
                    v2s.append(int(v[0]))
                    count += 1
                else:
                    v3s.append(int(v[0]))
    return v0s, v1s, np.array(v2s, np.int64), np.array(v3s, np.int64)


def flatten(vertices, edges, eps=1e-6):
    """flatten_loss.py:38-79: vertices [B,nv,3] -> [B]."""
    x = np.asarray(vertices, np.float64)
    v0s, v1s, v2s, v3s = (x[:, e, :] for e in edges)

    def half(b):
        a = v1s - v0s
        al2 = (a ** 2).sum(-1)
        bl2 = (b ** 2).sum(-1)
        al1 = np.sqrt(al2 + eps)
        bl1 = np.sqrt(bl2 + eps)
        ab = (a * b).sum(-1)
        cos = ab / (al1 * bl1 + eps)
        sin = np.sqrt(1 - cos ** 2 + eps)
        c = a * (ab / (al2 + eps))[..., None]
        return b - c, bl1 * sin
    cb1, cb1l1 = half(v2s - v0s)
    cb2, cb2l1 = half(v3s - v0s)
    cos = (cb1 * cb2).sum(-1) / (cb1l1 * cb2l1 + eps)
    return ((cos + 1) ** 2).sum(1)


def numeric_grad(fn, x, weights, h=1e-6):
    """d (sum_b weights[b] * fn(x)[b]) / d x by central differences in float64 (x [B,nv,3])."""
    x = np.asarray(x, np.float64)
    g = np.zeros_like(x)
    w = np.asarray(weights, np.float64)
    # every batch element's loss depends on its own vertices only: perturb the same coordinate of all of them at once
    for i in range(x.shape[1]):
        for k in range(3):
            xp, xm = x.copy(), x.copy()
            xp[:, i, k] += h
            xm[:, i, k] -= h
            g[:, i, k] = w * (fn(xp) - fn(xm)) / (2 * h)
    return g
