"""OBJ / MTL loading and texture baking (load-time, numpy).

Mirrors, for the SoftRas loader:
  jrender/io/load_obj.py:9-20
  jrender/io/utils/_load_obj_for_softras.py:19-207   (OBJ fan triangulation, MTL Kd / map_Kd)
  jrender/io/utils/load_textures.py:3-101            (bilinear bake of the image into R*R texels
                                                      per face -- a CUDA kernel in the reference,
                                                      restated here in numpy because it runs once
                                                      per mesh, not per frame)
Images are read with OpenCV (the reference uses skimage, which is not installed here).
Normal / bump maps and TBN (SSS, render2) are outside the hot-path scope.
"""
import os

import numpy as np
import torch


def load_mtl(filename_mtl):
    """_load_obj_for_softras.py:19-40: Kd colours and map_Kd filenames per material."""
    texture_filenames, colors, material_name = {}, {}, ''
    with open(filename_mtl) as f:
        for line in f.readlines():
            s = line.split()
            if len(s) == 0:
                continue
            if s[0] == 'newmtl':
                material_name = s[1]
            if s[0] == 'map_Kd':
                texture_filenames[material_name] = s[1]
            if s[0] == 'Kd':
                colors[material_name] = np.array(list(map(float, s[1:4])))
    return colors, texture_filenames


def _imread_rgb01(path):
    import cv2
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise IOError("cannot read texture image %s" % path)
    if img.ndim == 2:
        img = np.stack((img,) * 3, -1)
    else:
        img = img[:, :, :3][:, :, ::-1]   # BGR(A) -> RGB, drop alpha like the reference
    return img.astype(np.float32) / 255.


def bake_textures_for_softras(image, faces_uv, textures, is_update):
    """numpy restatement of load_textures_cuda_kernel (load_textures.py:11-69).

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


def load_textures(filename_obj, filename_mtl, texture_res):
    """_load_obj_for_softras.py:43-140 without the normal-map branch: [nf, R*R, 3] float32."""
    with open(filename_obj) as f:
        lines = f.readlines()
    vt = [[float(v) for v in line.split()[1:3]] for line in lines if len(line.split()) and line.split()[0] == 'vt']
    faces, material_names, material_name = [], [], ''
    for line in lines:
        s = line.split()
        if len(s) == 0:
            continue
        if s[0] == 'f':
            vs = s[1:]

            def ti(tok):
                return int(tok.split('/')[1]) if ('/' in tok and '//' not in tok) else 0
            v0 = ti(vs[0])
            for i in range(len(vs) - 2):
                faces.append((v0, ti(vs[i + 1]), ti(vs[i + 2])))
                material_names.append(material_name)
        if s[0] == 'usemtl':
            material_name = s[1]
    vt = np.vstack(vt).astype(np.float32)
    faces_uv = vt[np.vstack(faces).astype(np.int32) - 1]
    colors, texture_filenames = load_mtl(filename_mtl)
    textures = np.ones((faces_uv.shape[0], 3), np.float32)
    names = np.array(material_names)
    for material_name, color in colors.items():
        textures[names == material_name] = color
    textures = np.repeat(textures[:, None, :], texture_res ** 2, axis=1).astype(np.float32)
    for material_name, filename_texture in texture_filenames.items():
        image = _imread_rgb01(os.path.join(os.path.dirname(filename_obj), filename_texture))[::-1, :, :]
        textures = bake_textures_for_softras(np.ascontiguousarray(image), faces_uv, textures, (names == material_name).astype(np.int32))
    return textures


def load_obj(filename_obj, normalization=False, load_texture=False, dr_type='softras', texture_res=4,
             texture_type='surface', texture_wrapping='REPEAT', use_bilinear=True):
    """load_obj.py:9-20 / _load_obj_for_softras.py:142-207.  Returns torch tensors:
    (vertices [nv,3] f32, faces [nf,3] i32) or (+ textures) when load_texture."""
    assert dr_type in ['softras', 'n3mr']
    assert texture_type in ['surface', 'vertex']
    with open(filename_obj) as f:
        lines = f.readlines()
    vertices = np.vstack([[float(v) for v in line.split()[1:4]] for line in lines
                          if len(line.split()) and line.split()[0] == 'v']).astype(np.float32)
    faces = []
    for line in lines:
        s = line.split()
        if len(s) and s[0] == 'f':
            vs = s[1:]
            v0 = int(vs[0].split('/')[0])
            for i in range(len(vs) - 2):
                faces.append((v0, int(vs[i + 1].split('/')[0]), int(vs[i + 2].split('/')[0])))
    faces = np.vstack(faces).astype(np.int32) - 1
    textures = None
    if load_texture and texture_type == 'surface':
        if dr_type != 'softras':
            raise NotImplementedError("texture baking for n3mr (ts^3 texels) is not mirrored yet")
        for line in lines:
            if line.startswith('mtllib'):
                textures = load_textures(filename_obj, os.path.join(os.path.dirname(filename_obj), line.split()[1]), texture_res)
        if textures is None:
            raise Exception('Failed to load textures.')
    elif load_texture and texture_type == 'vertex':
        textures = np.vstack([[float(v) for v in line.split()[4:7]] for line in lines
                              if len(line.split()) and line.split()[0] == 'v']).astype(np.float32)
    if normalization:   # unit cube centred at zero (:199-203)
        vertices = vertices - vertices.min(0)
        vertices = vertices / np.abs(vertices).max()
        vertices = vertices * 2
        vertices = vertices - vertices.max(0) / 2
    v, f = torch.from_numpy(vertices.astype(np.float32)), torch.from_numpy(faces)
    if load_texture:
        return v, f, torch.from_numpy(textures)
    return v, f


def save_obj(filename, vertices, faces):
    """Geometry-only export (jrender/io/save_obj.py:31-84 without the texture atlas)."""
    v = vertices.detach().cpu().numpy() if isinstance(vertices, torch.Tensor) else np.asarray(vertices)
    f = faces.detach().cpu().numpy() if isinstance(faces, torch.Tensor) else np.asarray(faces)
    assert v.ndim == 2 and f.ndim == 2
    with open(filename, 'w') as fh:
        fh.write('# %s\n\n' % os.path.basename(filename))
        for p in v:
            fh.write('v %.8f %.8f %.8f\n' % (p[0], p[1], p[2]))
        fh.write('\n')
        for t in f:
            fh.write('f %d %d %d\n' % (t[0] + 1, t[1] + 1, t[2] + 1))
