"""CPU tests of the host-side mirror (PyTorch): cameras, lighting, mesh, losses, OBJ IO.
These classes sit either side of the rasterizer (SURVEY.md section 8f "next" rows); they
are checked against direct numpy restatements of the reference formulas."""
import math
import os

import numpy as np
import pytest
import torch

import jrender_b200 as jr
from jrender_b200 import workloads as wl


def test_look_at_perspective_match_numpy_restatement():
    v, f = wl.sphere_by_faces(280)
    vt = torch.from_numpy(v)[None].repeat(3, 1, 1)
    eyes = np.asarray([wl.get_points_from_angles(2.732, 30, a) for a in (0, 120, 240)], np.float32)
    a = jr.perspective(jr.look_at(vt, torch.from_numpy(eyes)), 30.).numpy()
    b = wl.perspective(wl.look_at(np.repeat(v[None], 3, 0), eyes), 30.)
    assert np.abs(a - b).max() < 1e-6
    assert np.allclose(jr.face_vertices(torch.from_numpy(a), torch.from_numpy(f)[None].repeat(3, 1, 1)).numpy(),
                       wl.face_vertices(b, f))


def test_get_points_from_angles_scalar_and_tensor():
    p = jr.get_points_from_angles(2.732, 30., 40.)
    q = jr.get_points_from_angles(torch.tensor([2.732]), torch.tensor([30.]), torch.tensor([40.]))
    assert np.allclose(np.asarray(p), q.numpy()[0], atol=1e-6)
    assert abs(math.sqrt(sum(x * x for x in p)) - 2.732) < 1e-6


def test_transform_errors_like_reference():
    with pytest.raises(ValueError):
        jr.Transform(camera_mode='fisheye')
    t = jr.Transform(camera_mode='projection', K=np.eye(3)[None], R=np.eye(3)[None], t=np.zeros((1, 1, 3)))
    with pytest.raises(ValueError):
        t.set_eyes_from_angles(1., 0., 0.)
    with pytest.raises(ValueError):
        jr.look_at(torch.zeros(4, 3), [0, 0, -1])


def test_renderer_defaults_and_modes():
    r = jr.Renderer()
    assert r.rasterizer.image_size == 256 and r.rasterizer.fill_back is True
    assert abs(r.transform.eyes[2] + (1. / math.tan(math.radians(30)) + 1)) < 1e-9
    r.set_sigma(3e-5); r.set_gamma(2e-4)
    assert r.rasterizer.sigma_val == 3e-5 and r.rasterizer.gamma_val == 2e-4
    with pytest.raises(ValueError):
        jr.Renderer(dr_type='raytrace')
    with pytest.raises(AssertionError):
        r.set_texture_mode('volume')


def test_lighting_surface_diffuse_only():
    v, f = wl.sphere_by_faces(280)
    mesh = jr.Mesh(v, f)
    mesh.with_specular = False
    lit = jr.Lighting()(mesh)
    fv = wl.face_vertices(v[None], f)[0].astype(np.float64)
    n = np.cross(fv[:, 2] - fv[:, 1], fv[:, 0] - fv[:, 1])
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    expect = np.clip(0.5 + 0.5 * np.maximum(n[:, 1], 0.0), 0, 1)          # ambient 0.5 + 0.5*relu(n.y)
    assert np.abs(lit.textures[0, :, 0, 0].numpy() - expect).max() < 1e-5
    # default (with_specular=True) adds a non-negative specular term and stays clamped
    mesh2 = jr.Mesh(v, f)
    lit2 = jr.Lighting()(mesh2, eyes=wl.get_points_from_angles(2.732, 30, 0))
    assert float(lit2.textures.max()) <= 1.0 and float(lit2.textures.min()) >= 0.0


def test_vertex_normals_point_outward_on_sphere():
    v, f = wl.sphere_by_faces(280)
    mesh = jr.Mesh(v, f, texture_type='vertex')
    n = mesh.vertex_normals[0].numpy()
    cosang = (n * (v / np.linalg.norm(v, axis=1, keepdims=True))).sum(1)
    assert cosang.min() > 0.95
    assert tuple(mesh.face_textures.shape) == (1, 280, 3, 3)


def test_losses_match_dense_reference_formulas():
    v, f = wl.sphere_by_faces(280)
    nv = v.shape[0]
    lap = np.zeros([nv, nv], np.float32)           # laplacian_loss.py:11-26
    lap[f[:, 0], f[:, 1]] = -1; lap[f[:, 1], f[:, 0]] = -1
    lap[f[:, 1], f[:, 2]] = -1; lap[f[:, 2], f[:, 1]] = -1
    lap[f[:, 2], f[:, 0]] = -1; lap[f[:, 0], f[:, 2]] = -1
    r, c = np.diag_indices(nv)
    lap[r, c] = -lap.sum(1)
    lap /= lap[r, c][:, None]
    x = (v + np.random.default_rng(0).normal(0, 0.01, v.shape)).astype(np.float32)
    expect = ((lap @ x) ** 2).sum()
    got = jr.LaplacianLoss(torch.from_numpy(v), torch.from_numpy(f))(torch.from_numpy(x)[None])
    assert abs(float(got) - expect) < 1e-4 * expect
    fl = jr.FlattenLoss(torch.from_numpy(f))
    assert fl.v0s.numel() == 3 * 280 // 2                       # closed manifold: E = 3F/2
    assert float(fl(torch.from_numpy(x)[None])) >= 0
    p = torch.rand(2, 8, 8); t = (torch.rand(2, 8, 8) > 0.5).float()
    i = (p * t).sum((1, 2)); u = (p + t - p * t).sum((1, 2)) + 1e-6
    assert abs(float(jr.neg_iou_loss(p, t)) - float(1 - (i / u).sum() / 2)) < 1e-6


def test_obj_roundtrip_and_texture_bake(tmp_path):
    v, f = wl.sphere_by_faces(280)
    p = os.path.join(tmp_path, "s.obj")
    jr.save_obj(p, v, f)
    v2, f2 = jr.load_obj(p)
    assert np.allclose(v2.numpy(), v, atol=1e-6) and np.array_equal(f2.numpy(), f)
    # texture bake: a constant image bakes to that constant; a horizontal ramp to the texel's u
    from oracle.bake import bake_textures_for_softras   # the CPU checker; the product bakes on the GPU (tests/test_bake_gpu.py)
    img = np.full((16, 32, 3), 0.25, np.float32)
    uv = np.random.default_rng(0).uniform(0.05, 0.9, (10, 3, 2)).astype(np.float32)
    out = bake_textures_for_softras(img, uv, np.ones((10, 9, 3), np.float32), np.ones(10, np.int32))
    assert np.allclose(out, 0.25, atol=1e-6)
    ramp = np.repeat(np.linspace(0, 1, 32, dtype=np.float32)[None, :, None], 16, 0).repeat(3, 2)
    out = bake_textures_for_softras(ramp, uv, np.ones((10, 9, 3), np.float32), np.ones(10, np.int32))
    u_centre = (uv[:, 0, 0] * (1 / 9) + uv[:, 1, 0] * (1 / 9) + uv[:, 2, 0] * (7 / 9))   # texel (0,0): w0=w1=1/9
    assert np.abs(out[:, 0, 0] - u_centre).max() < 2e-2
    keep = bake_textures_for_softras(ramp, uv, np.ones((10, 9, 3), np.float32), np.zeros(10, np.int32))
    assert np.all(keep == 1)
