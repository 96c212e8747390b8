"""CPU oracle of the pre-raster geometry stage (TEST INFRASTRUCTURE; product code never imports it).

float32 numpy restatement, expression by expression, of
    jrender/renderer/transform/look_at.py:24-38      look_at
    jrender/renderer/transform/look.py:23-53         look
    jrender/renderer/transform/perspective.py:11-16  perspective
    jrender/renderer/transform/orthogonal.py:12-15   orthogonal
    jrender/structures/utils/faces_vertices.py:14-19 face_vertices
and the analytic vector-Jacobian product of that chain (float64) used to check the fused backward.
Pinned against the reference's own Python run through oracle/jittor_numpy_stub.py
(tests/golden/ref_host_transform_*.npz, generator oracle/make_ref_host_golden.py).
"""
import math

import numpy as np

F = np.float32


def _normalize(v, eps=1e-5):
    n = np.sqrt((v * v).sum(-1, keepdims=True, dtype=F)).astype(F)
    return (v / np.maximum(n, F(eps))).astype(F)


def camera_axes(eye, camera_mode="look_at", at=(0, 0, 0), direction=(0, 0, 1), up=(0, 1, 0), coordinate="right"):
    """eye [B,3] -> r [B,3,3] (rows x, y, z axes).  look_at.py:26-32 / look.py:23-46."""
    eye = np.asarray(eye, F)
    B = eye.shape[0]
    up = np.broadcast_to(np.asarray(up, F), (B, 3))
    if camera_mode == "look_at":
        z = _normalize(np.broadcast_to(np.asarray(at, F), (B, 3)) - eye)
    else:
        z = np.broadcast_to(_normalize(np.asarray(direction, F)), (B, 3))
        up = np.broadcast_to(_normalize(np.asarray(up[0], F)), (B, 3))
    if camera_mode == "look" and coordinate == "left":
        x = _normalize(np.cross(z, up).astype(F))
        y = _normalize(np.cross(x, z).astype(F))
    else:
        x = _normalize(np.cross(up, z).astype(F))
        y = _normalize(np.cross(z, x).astype(F))
    return np.stack([x, y, z], axis=1).astype(F)


def to_camera(vertices, eye, r):
    """look_at.py:34-38: (v - eye) @ r^T, accumulated in k order like a plain matmul."""
    d = (np.asarray(vertices, F) - np.asarray(eye, F)[:, None, :]).astype(F)
    out = np.empty(d.shape[:2] + (3,), F)
    for i in range(3):
        acc = (d[:, :, 0] * r[:, None, i, 0]).astype(F)
        acc = (acc + (d[:, :, 1] * r[:, None, i, 1]).astype(F)).astype(F)
        acc = (acc + (d[:, :, 2] * r[:, None, i, 2]).astype(F)).astype(F)
        out[:, :, i] = acc
    return out


def perspective_width(angle):
    return F(np.tan(F(angle / 180 * math.pi)))


def project(cam, perspective=True, viewing_angle=30.0, viewing_scale=1.0):
    z = cam[:, :, 2]
    if perspective:
        w = perspective_width(viewing_angle)
        x = ((cam[:, :, 0] / z).astype(F) / w).astype(F)   # perspective.py:14
        y = ((cam[:, :, 1] / z).astype(F) / w).astype(F)   # :15
    else:
        x = (cam[:, :, 0] * F(viewing_scale)).astype(F)    # orthogonal.py:13-14
        y = (cam[:, :, 1] * F(viewing_scale)).astype(F)
    return np.stack([x, y, z], axis=2).astype(F)


def project_faces(vertices, faces, eye, camera_mode="look_at", at=(0, 0, 0), direction=(0, 0, 1), up=(0, 1, 0),
                  coordinate="right", perspective=True, viewing_angle=30.0, viewing_scale=1.0):
    """vertices [B|1,nv,3], faces [B|1,nf,3], eye [B|1,3] -> face_vertices [B,nf,3,3]."""
    vertices = np.asarray(vertices, F)
    faces = np.asarray(faces)
    eye = np.asarray(eye, F).reshape(-1, 3)
    B = max(vertices.shape[0], faces.shape[0], eye.shape[0])
    vertices = np.broadcast_to(vertices, (B,) + vertices.shape[1:])
    faces = np.broadcast_to(faces, (B,) + faces.shape[1:])
    eye = np.broadcast_to(eye, (B, 3))
    r = camera_axes(eye, camera_mode, at, direction, up, coordinate)
    pv = project(to_camera(vertices, eye, r), perspective, viewing_angle, viewing_scale)
    return np.stack([pv[b][faces[b].astype(np.int64)] for b in range(B)], axis=0).astype(F)   # faces_vertices.py:16-19


def project_faces_vjp(vertices, faces, eye, grad_face_vertices, vertices_batch=None, **kw):
    """float64 vector-Jacobian product: grad_face_vertices [B,nf,3,3] -> grad_vertices [Bv,nv,3]."""
    vertices = np.asarray(vertices, np.float64)
    faces = np.asarray(faces)
    eye = np.asarray(eye, np.float64).reshape(-1, 3)
    g = np.asarray(grad_face_vertices, np.float64)
    B = g.shape[0]
    Bv = vertices.shape[0] if vertices_batch is None else vertices_batch
    nv = vertices.shape[1]
    faces = np.broadcast_to(faces, (B,) + faces.shape[1:])
    eyeb = np.broadcast_to(eye, (B, 3))
    r = camera_axes(eyeb.astype(F), kw.get("camera_mode", "look_at"), kw.get("at", (0, 0, 0)), kw.get("direction", (0, 0, 1)),
                    kw.get("up", (0, 1, 0)), kw.get("coordinate", "right")).astype(np.float64)
    out = np.zeros((Bv, nv, 3), np.float64)
    persp = kw.get("perspective", True)
    w = float(perspective_width(kw.get("viewing_angle", 30.0))) if persp else float(kw.get("viewing_scale", 1.0))
    for b in range(B):
        vb = vertices[b if vertices.shape[0] > 1 else 0]
        cam = (vb - eyeb[b]) @ r[b].T
        gp = np.zeros((nv, 3), np.float64)
        np.add.at(gp, faces[b].reshape(-1).astype(np.int64), g[b].reshape(-1, 3))
        if persp:
            Z = cam[:, 2]
            dc = np.stack([gp[:, 0] / (Z * w), gp[:, 1] / (Z * w),
                           gp[:, 2] - (gp[:, 0] * cam[:, 0] + gp[:, 1] * cam[:, 1]) / (Z * Z * w)], axis=1)
        else:
            dc = np.stack([gp[:, 0] * w, gp[:, 1] * w, gp[:, 2]], axis=1)
        out[b if Bv > 1 else 0] += dc @ r[b]
    return out
