"""OBJ / MTL loading and texture baking (load-time, numpy).

Mirrors, for the SoftRas loader:
  jrender/io/load_obj.py:9-20
  jrender/io/utils/_load_obj_for_softras.py:19-207   (OBJ fan triangulation, MTL Kd / map_Kd)
  jrender/io/utils/load_textures.py:3-101            (bilinear bake of the image into R*R texels
                                                      per face -- a CUDA kernel in the reference,
                                                      restated here in numpy because it runs once
                                                      per mesh, not per frame)
and for the NMR loader (dr_type='n3mr'):
  jrender/io/utils/_load_obj_for_n3mr.py:11-155      (same parsing; [nf, ts, ts, ts, 3] textures, default colour 0.5)
  jrender/io/utils/load_textures.py:103-246          (the ts^3 bake with texture wrapping modes -> b200r_bake_textures_n3mr)
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


def bake_textures_for_softras(image, faces_uv, textures, is_update, device=None):
    """_load_textures_for_softras (load_textures.py:3-101) through the CUDA kernel b200r_bake_textures_softras.

    image [H,W,3] (already flipped vertically by the caller), faces_uv [nf,3,2], textures [nf,R*R,3], is_update [nf]:
    numpy arrays or tensors; returns a CUDA float32 tensor [nf,R*R,3] (updated where is_update != 0).
    There is no CPU implementation in the product (oracle/bake.py is the checker)."""
    import ctypes as C
    from . import _lib
    if not torch.cuda.is_available():
        raise _lib.B200RasterError("texture baking runs on the GPU (b200r_bake_textures_softras); there is no CPU fallback")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())

    def dv(x, dtype):
        t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
        return t.to(device=dev, dtype=dtype).contiguous()
    img, uv, upd = dv(image, torch.float32), dv(faces_uv, torch.float32), dv(is_update, torch.int32)
    out = dv(textures, torch.float32).clone()
    nf, T = out.shape[:2]
    R = int(round(float(np.sqrt(T))))
    if R * R != T or uv.shape != (nf, 3, 2) or img.dim() != 3 or img.shape[2] != 3 or upd.shape != (nf,):
        raise ValueError("bake_textures_for_softras: image [H,W,3], faces_uv [nf,3,2], textures [nf,R*R,3], is_update [nf]")
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b200r_bake_textures_softras(
            C.c_void_p(img.data_ptr()), C.c_void_p(uv.data_ptr()), C.c_void_p(upd.data_ptr()), C.c_void_p(out.data_ptr()),
            nf, R, img.shape[0], img.shape[1], C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "b200r_bake_textures_softras")
    return out


TEXTURE_WRAPPING = {'REPEAT': 0, 'MIRRORED_REPEAT': 1, 'CLAMP_TO_EDGE': 2, 'CLAMP_TO_BORDER': 3}   # _load_obj_for_n3mr.py:7-8


def bake_textures_for_n3mr(image, faces_uv, textures, is_update, texture_wrapping=0, use_bilinear=True, device=None):
    """_load_textures_for_n3mr (load_textures.py:103-246) through the CUDA kernel b200r_bake_textures_n3mr.

    image [H,W,3] (already flipped vertically by the caller), faces_uv [nf,3,2], textures [nf,ts,ts,ts,3], is_update [nf];
    texture_wrapping 0..3 (TEXTURE_WRAPPING) or its name.  Returns a CUDA float32 tensor [nf,ts,ts,ts,3], updated where
    is_update != 0.  No CPU implementation in the product (oracle/bake.py is the checker)."""
    import ctypes as C
    from . import _lib
    if not torch.cuda.is_available():
        raise _lib.B200RasterError("texture baking runs on the GPU (b200r_bake_textures_n3mr); there is no CPU fallback")
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    wrap = TEXTURE_WRAPPING[texture_wrapping] if isinstance(texture_wrapping, str) else int(texture_wrapping)

    def dv(x, dtype):
        t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
        return t.to(device=dev, dtype=dtype).contiguous()
    img, uv, upd = dv(image, torch.float32), dv(faces_uv, torch.float32), dv(is_update, torch.int32)
    out = dv(textures, torch.float32).clone()
    if out.dim() != 5 or out.shape[1] != out.shape[2] or out.shape[1] != out.shape[3] or out.shape[4] != 3:
        raise ValueError("bake_textures_for_n3mr: textures must be [nf,ts,ts,ts,3]")
    nf, ts = out.shape[0], out.shape[1]
    if uv.shape != (nf, 3, 2) or img.dim() != 3 or img.shape[2] != 3 or upd.shape != (nf,):
        raise ValueError("bake_textures_for_n3mr: image [H,W,3], faces_uv [nf,3,2], textures [nf,ts,ts,ts,3], is_update [nf]")
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b200r_bake_textures_n3mr(
            C.c_void_p(img.data_ptr()), C.c_void_p(uv.data_ptr()), C.c_void_p(upd.data_ptr()), C.c_void_p(out.data_ptr()),
            nf, ts, img.shape[0], img.shape[1], wrap, 1 if use_bilinear else 0,
            C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "b200r_bake_textures_n3mr")
    return out


def _parse_texture_faces(filename_obj):
    """UV triangles (fan triangulation, 0 for a missing vt index) and the material of every face
    (_load_obj_for_softras.py:43-83 == _load_obj_for_n3mr.py:30-68)."""
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
    return faces_uv, material_names


def load_textures_n3mr(filename_obj, filename_mtl, texture_res, texture_wrapping='REPEAT', use_bilinear=True):
    """_load_obj_for_n3mr.py:29-103: [nf, ts, ts, ts, 3] float32; faces without a material keep 0.5, Kd colours are
    broadcast over the ts^3 texels, map_Kd images are baked by b200r_bake_textures_n3mr."""
    faces_uv, material_names = _parse_texture_faces(filename_obj)
    colors, texture_filenames = load_mtl(filename_mtl)
    textures = np.full((faces_uv.shape[0], 3), 0.5, np.float32)
    names = np.array(material_names)
    for material_name, color in colors.items():
        textures[names == material_name] = color
    ts = int(texture_res)
    textures = np.ascontiguousarray(np.broadcast_to(textures[:, None, None, None, :], (faces_uv.shape[0], ts, ts, ts, 3)), dtype=np.float32)
    if ts < 2 and texture_filenames:
        # The reference's kernel divides by (texture_size - 1.): with Mesh.from_obj's default texture_res = 1 every sample
        # point is 0 / 0 and the baked texels are garbage (demo4-optim_textures.py loads the cow this way and then replaces
        # the textures).  Here the material colours are kept and the image is not sampled.
        import warnings
        warnings.warn("load_obj(dr_type='n3mr', texture_res=1): one texel per face cannot be sampled from the texture image "
                      "(the reference divides by texture_res - 1); material colours kept, use texture_res >= 2 to bake")
        return textures
    for material_name, filename_texture in texture_filenames.items():
        image = _imread_rgb01(os.path.join(os.path.dirname(filename_obj), filename_texture))[::-1, :, :]
        textures = bake_textures_for_n3mr(np.ascontiguousarray(image), faces_uv, textures, (names == material_name).astype(np.int32),
                                          TEXTURE_WRAPPING[texture_wrapping], use_bilinear)
    if isinstance(textures, torch.Tensor):
        textures = textures.cpu().numpy()
    return textures


def load_textures(filename_obj, filename_mtl, texture_res):
    """_load_obj_for_softras.py:43-140 without the normal-map branch: [nf, R*R, 3] float32."""
    faces_uv, material_names = _parse_texture_faces(filename_obj)
    colors, texture_filenames = load_mtl(filename_mtl)
    textures = np.ones((faces_uv.shape[0], 3), np.float32)
    names = np.array(material_names)
    for material_name, color in colors.items():
        textures[names == material_name] = color
    textures = np.repeat(textures[:, None, :], texture_res ** 2, axis=1).astype(np.float32)
    for material_name, filename_texture in texture_filenames.items():
        image = _imread_rgb01(os.path.join(os.path.dirname(filename_obj), filename_texture))[::-1, :, :]
        textures = bake_textures_for_softras(np.ascontiguousarray(image), faces_uv, textures, (names == material_name).astype(np.int32))
    if isinstance(textures, torch.Tensor):   # baked on the GPU; loaders hand back host arrays like the geometry
        textures = textures.cpu().numpy()
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
    if load_texture and dr_type == 'n3mr':   # _load_obj_for_n3mr.py:139-147 (no texture_type there)
        for line in lines:
            if line.startswith('mtllib'):
                textures = load_textures_n3mr(filename_obj, os.path.join(os.path.dirname(filename_obj), line.split()[1]), texture_res,
                                              texture_wrapping=texture_wrapping, use_bilinear=use_bilinear)
        if textures is None:
            raise Exception('Failed to load textures.')
    elif load_texture and texture_type == 'surface':
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
