"""CPU tests of the NMR oracle (oracle/nmr_oracle.c): invariants that follow from the reference
source (jrender/renderer/dr/n3mr/cuda/rasterize.py), independent of any GPU."""
import numpy as np

from oracle import nmr as onmr
from tests.util import nmr_scene


def test_forward_invariants():
    faces, tex = nmr_scene(280, batch=2, ts=2)
    out = onmr.forward(faces, tex, 64, 0.1, 100.0, 1e-3, (0.1, 0.2, 0.3))
    idx = out["face_index_map"]
    cov = idx >= 0
    assert 0.1 < cov.mean() < 0.5
    assert np.all(out["depth_map"][~cov] == np.float32(100.0))            # thrust::fill(far) :181-182
    assert np.all(out["weight_map"][~cov] == 0) and np.all(out["alpha_map"] == cov)
    w = out["weight_map"][cov]
    assert np.abs(w.sum(-1) - 1).max() < 1e-5 and w.min() >= 0             # clamped + normalised :126-133
    assert np.allclose(out["rgb_map"][~cov], [0.1, 0.2, 0.3])              # background mix n3mr.py:135-143
    assert np.abs(out["sampling_weight_map"][cov].sum(-1) - 1).max() < 1e-5  # trilinear weights sum to 1
    assert out["sampling_index_map"].max() < 8 and out["sampling_index_map"].min() >= 0
    # fill_back: the reversed copies are back-facing exactly where the originals are front-facing
    assert idx.max() < faces.shape[1]


def test_lowest_id_wins_equal_depth():
    tri = np.array([[[-0.5, -0.5, 2.0], [0.5, -0.5, 2.0], [0.0, 0.6, 2.0]]], np.float32)
    faces = np.stack([tri, tri], 1).reshape(1, 2, 3, 3)                    # two identical faces
    out = onmr.forward(faces, None, 32, 0.1, 100.0, 1e-3, (0, 0, 0), False, True, False)
    assert set(np.unique(out["face_index_map"])) == {-1, 0}


def test_backward_shapes_and_culling():
    faces, tex = nmr_scene(280, batch=1, ts=2)
    out = onmr.forward(faces, tex, 48)
    rng = np.random.default_rng(0)
    g = rng.uniform(-1, 1, (1, 48, 48, 3)).astype(np.float32)
    ga = rng.uniform(-1, 1, (1, 48, 48)).astype(np.float32)
    gd = rng.uniform(-1, 1, (1, 48, 48)).astype(np.float32)
    gf, gt = onmr.backward(faces, tex, out, 48, 1e-3, g, ga, gd)
    assert gf.shape == faces.shape and gt.shape == tex.shape and not np.isnan(gf).any()
    never_seen = np.setdiff1d(np.arange(faces.shape[1]), np.unique(out["face_index_map"]))
    assert np.all(gt[0, never_seen] == 0)                                   # texture grads only for visible faces
    gf2, _ = onmr.backward(faces, tex, out, 48, 1e-3, g, ga, gd, True, True, False)
    assert np.all(gf2[..., 2] == 0) and np.abs(gf[..., 2]).max() > 0        # z gradient comes from the depth term only
