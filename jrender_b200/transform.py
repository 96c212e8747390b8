"""Camera transforms (host-side mirror in PyTorch; no custom kernels -- SURVEY.md section 2.1 "next").

Mirrors jrender/renderer/transform/{look_at,look,perspective,orthogonal,projection,transform}.py
and jrender/renderer/utils/get_points_from_angles.py: same names, arguments, defaults, errors.
"""
import math

import numpy as np
import torch
from torch import nn
import torch.nn.functional as F


_CONST = {}


def _t(x, like=None, dtype=torch.float32):
    """Tensors pass through; Python / numpy constants (at, up, direction, a fixed eye) are placed on
    the device once per (value, device) -- a per-call H2D copy would stall the stream and break CUDA
    graph capture."""
    if isinstance(x, torch.Tensor):
        t = x.to(dtype)
        return t.to(like.device) if like is not None else t
    a = np.asarray(x, dtype=np.float32)
    key = (a.tobytes(), a.shape, str(dtype), str(like.device) if like is not None else "cpu")
    t = _CONST.get(key)
    if t is None:
        t = torch.tensor(a, dtype=dtype)
        if like is not None:
            t = t.to(like.device)
        if len(_CONST) < 4096:
            _CONST[key] = t
    return t


def get_points_from_angles(distance, elevation, azimuth, degrees=True):
    """utils/get_points_from_angles.py:4-22."""
    if isinstance(distance, (float, int)):
        if degrees:
            elevation = math.radians(elevation)
            azimuth = math.radians(azimuth)
        return (distance * math.cos(elevation) * math.sin(azimuth),
                distance * math.sin(elevation),
                -distance * math.cos(elevation) * math.cos(azimuth))
    if degrees:
        elevation = math.pi / 180. * elevation
        azimuth = math.pi / 180. * azimuth
    return torch.stack([distance * torch.cos(elevation) * torch.sin(azimuth),
                        distance * torch.sin(elevation),
                        -distance * torch.cos(elevation) * torch.cos(azimuth)], dim=0).transpose(1, 0)


def _normalize(v, eps=1e-5, dim=-1):
    return F.normalize(v, p=2, dim=dim, eps=eps)


def look_at(vertices, eye, at=[0, 0, 0], up=[0, 1, 0]):
    """look_at.py:3-39."""
    if vertices.dim() != 3:
        raise ValueError('vertices Tensor should have 3 dimensions')
    at = _t(at, vertices)
    up = _t(up, vertices)
    eye = _t(list(eye) if isinstance(eye, tuple) else eye, vertices)
    batch_size = vertices.shape[0]
    if eye.dim() == 1:
        eye = eye[None, :].expand(batch_size, -1)
    if at.dim() == 1:
        at = at[None, :].expand(batch_size, -1)
    if up.dim() == 1:
        up = up[None, :].expand(batch_size, -1)
    z_axis = _normalize(at - eye)
    x_axis = _normalize(torch.cross(up, z_axis, dim=-1))
    y_axis = _normalize(torch.cross(z_axis, x_axis, dim=-1))
    r = torch.cat((x_axis[:, None, :], y_axis[:, None, :], z_axis[:, None, :]), dim=1)
    if vertices.shape != eye.shape:
        eye = eye[:, None, :]
    vertices = vertices - eye
    return torch.matmul(vertices, r.transpose(1, 2))


def look(vertices, eye, direction=[0, 1, 0], up=None, coordinate="right"):
    """look.py:3-54."""
    if vertices.dim() != 3:
        raise ValueError('vertices Tensor should have 3 dimensions')
    direction = _t(direction, vertices)
    eye = _t(list(eye) if isinstance(eye, tuple) else eye, vertices)
    up = _t([0, 1, 0] if up is None else up, vertices)
    z_axis = _normalize(direction, dim=0)
    up = _normalize(up, dim=0)
    if torch.abs(torch.sum(up * z_axis)) > 1 - 1e-4:
        raise ValueError("camera_direction and camera_up can not be the same")
    batch_size = vertices.shape[0]
    if eye.dim() == 1:
        eye = eye[None, :].expand(batch_size, -1)
    if z_axis.dim() == 1:
        z_axis = z_axis[None, :].expand(batch_size, -1)
    if up.dim() == 1:
        up = up[None, :].expand(batch_size, -1)
    if coordinate == "right":
        x_axis = _normalize(torch.cross(up, z_axis, dim=-1))
        y_axis = _normalize(torch.cross(z_axis, x_axis, dim=-1))
    elif coordinate == "left":
        x_axis = _normalize(torch.cross(z_axis, up, dim=-1))
        y_axis = _normalize(torch.cross(x_axis, z_axis, dim=-1))
    else:
        raise ValueError("coordinate must be 'right' or 'left'")
    r = torch.cat((x_axis[:, None, :], y_axis[:, None, :], z_axis[:, None, :]), dim=1)
    if vertices.shape != eye.shape:
        eye = eye[:, None, :]
    vertices = vertices - eye
    return torch.matmul(vertices, r.transpose(1, 2))


_TAN_HALF_ANGLE = {}


def perspective(vertices, angle=30.):
    """perspective.py:4-17."""
    if vertices.dim() != 3:
        raise ValueError('vertices Tensor should have 3 dimensions')
    key = (float(angle), str(vertices.device))
    width = _TAN_HALF_ANGLE.get(key)   # device constant, built once: no per-call H2D copy (CUDA-graph safe)
    if width is None:
        width = _TAN_HALF_ANGLE[key] = torch.tan(
            torch.tensor([angle / 180 * math.pi], dtype=torch.float32, device=vertices.device))[:, None]
    z = vertices[:, :, 2]
    x = vertices[:, :, 0] / z / width
    y = vertices[:, :, 1] / z / width
    return torch.stack((x, y, z), dim=2)


def orthogonal(vertices, scale):
    """orthogonal.py:3-15."""
    if vertices.dim() != 3:
        raise ValueError('vertices Tensor should have 3 dimensions')
    z = vertices[:, :, 2]
    x = vertices[:, :, 0] * scale
    y = vertices[:, :, 1] * scale
    return torch.stack((x, y, z), dim=2)


def projection(vertices, K, R, t, dist_coeffs, orig_size, eps=1e-9):
    """projection.py:3-48 (note: uses R[0] / K[0] for the whole batch, like the reference)."""
    vertices = torch.matmul(vertices, R.transpose(1, 2)[0]) + t
    x, y, z = vertices[:, :, 0], vertices[:, :, 1], vertices[:, :, 2]
    x_ = x / (z + eps)
    y_ = y / (z + eps)
    k1 = dist_coeffs[:, 0][:, None]
    k2 = dist_coeffs[:, 1][:, None]
    p1 = dist_coeffs[:, 2][:, None]
    p2 = dist_coeffs[:, 3][:, None]
    k3 = dist_coeffs[:, 4][:, None]
    x_2 = x_ * x_
    y_2 = y_ * y_
    r = torch.sqrt(x_2 + y_2)
    r2 = r * r
    r4 = r2 * r2
    r6 = r4 * r2
    tmp = k1 * r2 + k2 * r4 + k3 * r6 + 1
    x__ = x_ * tmp + 2 * p1 * x_ * y_ + p2 * (r2 + 2 * x_2)
    y__ = y_ * tmp + p1 * (r2 + 2 * y_2) + 2 * p2 * x_ * y_
    vertices = torch.stack([x__, y__, torch.ones_like(z)], dim=-1)
    vertices = torch.matmul(vertices, K.transpose(1, 2)[0])
    u, v = vertices[:, :, 0], vertices[:, :, 1]
    v = orig_size - v
    u = 2 * (u - orig_size / 2.) / orig_size
    v = 2 * (v - orig_size / 2.) / orig_size
    return torch.stack([u, v, z], dim=-1)


class Projection(nn.Module):
    def __init__(self, K, R, t, dist_coeffs=None, orig_size=512):
        super(Projection, self).__init__()
        self.K = _t(K) if isinstance(K, np.ndarray) else K
        self.R = _t(R) if isinstance(R, np.ndarray) else R
        self.t = _t(t) if isinstance(t, np.ndarray) else t
        self.dist_coeffs = dist_coeffs
        self.orig_size = orig_size
        self._eye = None
        if dist_coeffs is None:
            self.dist_coeffs = torch.zeros((self.K.shape[0], 5), dtype=torch.float32)

    def forward(self, vertices):
        d = vertices.device
        return projection(vertices, self.K.to(d), self.R.to(d), self.t.to(d), self.dist_coeffs.to(d), self.orig_size)

    execute = forward


class LookAt(nn.Module):
    def __init__(self, perspective=True, viewing_angle=30, viewing_scale=1.0, eye=None):
        super(LookAt, self).__init__()
        self.perspective = perspective
        self.viewing_angle = viewing_angle
        self.viewing_scale = viewing_scale
        self._eye = eye
        if self._eye is None:
            self._eye = [0, 0, -(1. / math.tan(math.radians(self.viewing_angle)) + 1)]

    def forward(self, vertices):
        vertices = look_at(vertices, self._eye)
        if self.perspective:
            vertices = perspective(vertices, angle=self.viewing_angle)
        else:
            vertices = orthogonal(vertices, scale=self.viewing_scale)
        return vertices

    execute = forward


class Look(nn.Module):
    def __init__(self, camera_direction=[0, 0, 1], perspective=True, viewing_angle=30, viewing_scale=1.0, eye=None,
                 up=[0, 1, 0], coordinate="right"):
        super(Look, self).__init__()
        self.perspective = perspective
        self.viewing_angle = viewing_angle
        self.viewing_scale = viewing_scale
        self._eye = eye
        self.camera_direction = camera_direction
        self.up = up
        self.coordinate = coordinate
        if self._eye is None:
            self._eye = [0, 0, -(1. / math.tan(math.radians(self.viewing_angle)) + 1)]

    def forward(self, vertices):
        vertices = look(vertices, self._eye, self.camera_direction, up=self.up, coordinate=self.coordinate)
        if self.perspective:
            vertices = perspective(vertices, angle=self.viewing_angle)
        else:
            vertices = orthogonal(vertices, scale=self.viewing_scale)
        return vertices

    execute = forward


class Transform(nn.Module):
    """transform.py:83-135."""

    def __init__(self, camera_mode='projection', K=None, R=None, t=None, dist_coeffs=None, orig_size=512,
                 perspective=True, viewing_angle=30, viewing_scale=1.0,
                 eye=None, camera_direction=[0, 0, 1], up=[0, 1, 0], coordinate="right"):
        super(Transform, self).__init__()
        self.camera_mode = camera_mode
        if self.camera_mode == 'projection':
            self.transformer = Projection(K, R, t, dist_coeffs, orig_size)
        elif self.camera_mode == 'look':
            self.transformer = Look(camera_direction, perspective, viewing_angle, viewing_scale, eye, up, coordinate)
        elif self.camera_mode == 'look_at':
            self.transformer = LookAt(perspective, viewing_angle, viewing_scale, eye)
        else:
            raise ValueError('Camera mode has to be one of projection, look or look_at')
        self.eye = eye
        self.camera_direction = camera_direction
        self.viewing_angle = viewing_angle
        self.up = up
        self.coordinate = coordinate

    def forward(self, mesh):
        proj = self._fused_projector(mesh)
        if proj is not None:
            # One fused launch (csrc/preraster_api.cu) produces the rasterizer's face_vertices straight from
            # the world-space vertices; `mesh.vertices` (camera space) stays available, evaluated lazily by
            # the op-by-op mirror only if somebody reads it.
            mesh.attach_projector(proj)
        else:
            mesh.vertices = self.transformer(mesh.vertices)
        return mesh

    fused = True  # class-wide switch: False forces the op-by-op PyTorch mirror

    def _fused_projector(self, mesh):
        v = mesh.vertices
        tr = self.transformer
        if not (self.fused and self.camera_mode in ('look_at', 'look') and v.is_cuda and v.dtype == torch.float32):
            return None
        eye = tr._eye
        if isinstance(eye, torch.Tensor) and eye.requires_grad:
            return None   # camera optimisation: gradients w.r.t. the eye come from the op-by-op mirror
        if not isinstance(eye, torch.Tensor) and isinstance(eye, (list, tuple)) and any(isinstance(c, torch.Tensor) for c in eye):
            return None
        from .preraster import project_faces
        kw = dict(camera_mode=self.camera_mode, perspective=tr.perspective, viewing_angle=tr.viewing_angle,
                  viewing_scale=tr.viewing_scale)
        if self.camera_mode == 'look':
            kw.update(direction=tr.camera_direction, up=tr.up, coordinate=tr.coordinate)
        transformer = tr

        class _Projector(object):
            def face_vertices(self, faces):
                return project_faces(v, faces, eye, **kw)

            def vertices(self):
                return transformer(v)
        return _Projector()

    execute = forward

    def tranpos(self, pos):
        return self.transformer(pos)

    def set_eyes_from_angles(self, distances, elevations, azimuths):
        if self.camera_mode not in ['look', 'look_at']:
            raise ValueError('Projection does not need to set eyes')
        self.transformer._eye = get_points_from_angles(distances, elevations, azimuths)

    def set_eyes(self, eyes):
        if self.camera_mode not in ['look', 'look_at']:
            raise ValueError('Projection does not need to set eyes')
        self.transformer._eye = eyes

    def view_transform(self, vertices):
        if self.camera_mode == 'look_at':
            vertices = look_at(vertices, self.eye)
        elif self.camera_mode == 'look':
            vertices = look(vertices, self.eye, self.camera_direction, up=self.up, coordinate=self.coordinate)
        return vertices

    def projection_transform(self, vertices):
        return perspective(vertices, self.viewing_angle)

    @property
    def eyes(self):
        return self.transformer._eye
