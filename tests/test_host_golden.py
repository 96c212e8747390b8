"""Host-mirror parity against the reference's OWN Python (CPU): losses and lighting kernels of
jrender_b200 vs fixtures produced by executing jrender/loss/*.py and
jrender/renderer/lighting/{ambient,directional}_lighting.py through the numpy jittor stub
(oracle/make_ref_host_golden.py, tests/golden/ref_host_{loss,lighting}*.npz)."""
import os

import numpy as np
import torch

import jrender_b200 as jr
from jrender_b200 import lighting as jl

G = os.path.join(os.path.dirname(__file__), "golden")


def _close(a, b, tol):
    b = np.asarray(b, np.float64)
    return np.abs(np.asarray(a, np.float64) - b).max() <= tol * max(np.abs(b).max(), 1e-30)


def test_losses_match_reference_python():
    g = np.load(os.path.join(G, "ref_host_loss_sphere280.npz"))
    v = torch.from_numpy(g["vertices"])
    f = torch.from_numpy(g["faces"])
    lap = jr.LaplacianLoss(v[0], f)(v).numpy()
    assert lap.shape == g["laplacian"].shape and _close(lap, g["laplacian"], 2e-5)   # (L x)^2 summed over 142 x 3 terms
    flat = jr.FlattenLoss(f)(v).numpy()
    assert flat.shape == g["flatten"].shape and _close(flat, g["flatten"], 2e-5)
    iou = float(jr.neg_iou_loss(torch.from_numpy(g["iou_predict"]), torch.from_numpy(g["iou_target"])))
    assert abs(iou - float(g["neg_iou"])) <= 1e-6


def test_lighting_kernels_match_reference_python():
    g = np.load(os.path.join(G, "ref_host_lighting.npz"))
    n = torch.from_numpy(g["normals"])
    for tag, spec in (("diffuse", False), ("specular", True)):
        d0 = jl.ambient_lighting(torch.zeros_like(n), 0.4, (1, 0.95, 0.9))
        d, s = jl.directional_lighting(d0, torch.zeros_like(n), n, 0.6, tuple(g["light_color"].tolist()),
                                       tuple(g["light_direction"].tolist()), torch.from_numpy(g["positions"]),
                                       torch.from_numpy(g["eye"]), spec, torch.from_numpy(g["metallic"]),
                                       torch.from_numpy(g["roughness"]))
        assert _close(d.numpy(), g["diffuse_" + tag], 5e-6), tag
        assert _close(s.numpy(), g["specular_" + tag], 5e-5), tag   # GGX ratio: (1 - cos)^5 and a2/denom^2 amplify ulps


def test_mesh_normals_match_reference_python():
    """Mesh.surface_normals (float64 cross, normalised, back to float32) and Mesh.vertex_normals (face normals
    scatter-added to the vertices) against structures/mesh.py:213-248 executed through the stub."""
    g = np.load(os.path.join(G, "ref_host_mesh_normals.npz"))
    mesh = jr.Mesh(torch.from_numpy(g["vertices"]), torch.from_numpy(g["faces"]))
    assert _close(mesh.surface_normals.numpy(), g["surface_normals"], 2e-6)
    assert _close(mesh.vertex_normals.numpy(), g["vertex_normals"], 2e-6)


def test_obj_loader_matches_reference_python():
    """jr.load_obj (geometry, fan triangulation of polygons, v/vt/vn index forms, unit-cube normalisation, per-vertex
    colours) against io/utils/_load_obj_for_softras.py:142-207 executed through the stub on the same OBJ text."""
    g = np.load(os.path.join(G, "ref_host_obj_loader.npz"))
    fn = os.path.join(G, "ref_host_obj_fixture.obj")
    v, f = jr.load_obj(fn)
    assert np.array_equal(v.numpy(), g["vertices_raw"])
    assert np.array_equal(f.numpy().astype(np.float32), g["faces"])          # the reference keeps faces as float32
    v1, f1 = jr.load_obj(fn, normalization=True)
    assert _close(v1.numpy(), g["vertices_normalized"], 1e-6)
    assert np.array_equal(f1.numpy().astype(np.float32), g["faces_normalized_call"])
    v2, f2, t2 = jr.load_obj(fn, normalization=True, load_texture=True, texture_type='vertex')
    assert np.array_equal(t2.numpy(), g["vertex_textures"]) and _close(v2.numpy(), g["vertices_normalized"], 1e-6)


def test_projection_camera_and_angle_points_match_reference_python():
    """jr.projection (camera_mode='projection': R/t, radial + tangential distortion, K, pixel -> NDC) and the tensor
    branch of jr.get_points_from_angles against transform/projection.py and utils/get_points_from_angles.py."""
    g = np.load(os.path.join(G, "ref_host_camera_extras.npz"))
    t = lambda k: torch.from_numpy(g[k])
    out = jr.projection(t("vertices"), t("K"), t("R"), t("t"), t("dist_coeffs"), int(g["orig_size"]))
    assert _close(out.numpy(), g["projected"], 5e-6)
    pts = jr.get_points_from_angles(t("distance"), t("elevation"), t("azimuth"))
    assert _close(pts.numpy(), g["points"], 2e-6)


def test_lighting_stage_matches_reference_python():
    """jr.Lighting (surface mode: ambient + directional, Cook-Torrance branch and diffuse-only) applied to a jr.Mesh
    against the reference's Lighting.execute (lighting/lighting.py:159-223) run through the stub on the same mesh."""
    g = np.load(os.path.join(G, "ref_host_lighting_stage.npz"))
    for tag, spec in (("specular", True), ("diffuse", False)):
        mesh = jr.Mesh(torch.from_numpy(g["vertices"]), torch.from_numpy(g["faces"]), textures=torch.from_numpy(g["textures"].copy()))
        mesh.with_specular = spec
        out = jr.Lighting()(mesh, torch.from_numpy(g["eyes"]))
        assert _close(out.textures.numpy(), g["lit_" + tag], 2e-5), tag
