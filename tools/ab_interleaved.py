"""Interleaved A/B of library builds inside ONE process (same GPU, same clocks): every round runs each
library's forward+backward a few times with its event profiler on; medians over rounds are printed.

    python tools/ab_interleaved.py c3|c5 lib1.so lib2.so ...      (raw C ABI through ctypes, no autograd)
"""
import ctypes as C
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jrender_b200 import workloads as wl  # noqa: E402

_F, _I, _P = C.c_float, C.c_int, C.c_void_p
SCAL = [_I, _I, _I, _I, _I, _F, _F, _F, _F, _F, _F, _I, _I, _I, _I, _I, _P]


def load(spec):
    """spec = path[#option=value[,option=value...]]: options are applied to this library instance."""
    path, _, opts = spec.partition("#")
    L = C.CDLL(path)
    L.b200r_set_option.argtypes = [C.c_char_p, _I]
    for kv in filter(None, opts.split(",")):
        k, v = kv.split("=")
        assert L.b200r_set_option(k.encode(), int(v)) == 0, kv
    L.b200r_softras_workspace_bytes.restype = C.c_size_t
    L.b200r_softras_workspace_bytes.argtypes = [_I, _I, _I]
    L.b200r_softras_forward.restype = _I
    L.b200r_softras_forward.argtypes = [_P] * 7 + [C.c_size_t] + SCAL
    L.b200r_softras_backward.restype = _I
    L.b200r_softras_backward.argtypes = [_P] * 6 + [C.c_size_t, _P, _P, _P] + SCAL
    L.b200r_profile_read.argtypes = [_I, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
    L.b200r_last_error.restype = C.c_char_p
    return L


def scene(name):
    if name == "c3":
        fv, tex = wl.make_scene(39200, batch=4)
        return fv, tex, 1024, 1e-5, 1   # rgb softmax
    if name == "c2":
        fv, tex = wl.make_scene(3280, batch=8)
        return fv, tex, 1024, 1e-5, 1
    views = 60
    v, f = wl.sphere_by_faces(3280, radius=1.0)
    v = v * 0.5
    eyes = np.asarray([wl.get_points_from_angles(5.464, 30.0 * math.sin(b), 360.0 * b / views) for b in range(views)], np.float32)
    fv = wl.face_vertices(wl.perspective(wl.look_at(np.repeat(v[None], views, 0), eyes), 15.0), f)
    return fv, np.ones((views, f.shape[0], 1, 3), np.float32), 512, 1e-4, 2   # silhouette: rgb none


def main():
    name, libs = sys.argv[1], sys.argv[2:]
    fv_h, tex_h, H, sigma, rgb = scene(name)
    dev = torch.device("cuda:0")
    B, nf = fv_h.shape[:2]
    K = 16
    fv, tex = torch.from_numpy(fv_h).to(dev), torch.from_numpy(tex_h).to(dev)
    out = torch.empty((B, 4, H, H), device=dev); aggr = torch.empty((B, 2, H, H), device=dev)
    ids = torch.empty((B, K, H, H), dtype=torch.int32, device=dev)
    g = torch.rand((B, 4, H, H), device=dev) * 2 - 1
    if rgb == 2:
        g[:, :3] = 0
    gf, gt = torch.empty_like(fv), torch.empty_like(tex)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    scal = (B, nf, 1, H, K, 1.0, 100.0, 1e-3, float(np.float32(sigma)), float(np.float32(1e-4)),
            float(np.float32(math.log(1.0 / 1e-4 - 1.0))), 2, rgb, 2, 0, 1)
    Ls = [(os.path.basename(l), load(l)) for l in libs]
    if len(set(l.partition('#')[0] for l in libs)) != len(libs):   # the same file loaded twice would share one handle (and its options)
        import shutil, tempfile
        tmp = tempfile.mkdtemp()
        Ls = []
        for i, l in enumerate(libs):
            path, _, opts = l.partition("#")
            cp = os.path.join(tmp, "%d_%s" % (i, os.path.basename(path)))
            shutil.copy(path, cp)
            Ls.append((os.path.basename(l), load(cp + ("#" + opts if opts else ""))))
    ws = {}
    for n_, L in Ls:
        nb = L.b200r_softras_workspace_bytes(B, nf, H)
        ws[n_] = (torch.empty(nb, dtype=torch.uint8, device=dev), nb)

    def step(n_, L):
        w, nb = ws[n_]
        rc = L.b200r_softras_forward(p(fv), p(tex), p(out), p(aggr), p(ids), None, p(w), nb, *scal, st)
        assert rc == 0, L.b200r_last_error()
        rc = L.b200r_softras_backward(p(fv), p(tex), p(out), p(aggr), p(ids), p(w), nb, p(g), p(gf), p(gt), *scal, st)
        assert rc == 0, L.b200r_last_error()
    for n_, L in Ls:
        for _ in range(3):
            step(n_, L)
    torch.cuda.synchronize()
    res = {n_: {"fwd": [], "bwd": []} for n_, _ in Ls}
    rounds, per = 7, 4
    for r in range(rounds):
        order = Ls if r % 2 == 0 else Ls[::-1]
        for n_, L in order:
            L.b200r_profile_reset(); L.b200r_profile_enable(1)
            for _ in range(per):
                flush.fill_(1.0)
                step(n_, L)
            torch.cuda.synchronize(); L.b200r_profile_enable(0)
            for kid, key in ((2, "fwd"), (3, "bwd")):
                ms, cnt = C.c_double(0), C.c_longlong(0)
                L.b200r_profile_read(kid, C.byref(ms), C.byref(cnt))
                res[n_][key].append(ms.value / max(1, cnt.value))
    print(json.dumps({"workload": name, **{n_: {k: [round(float(np.min(v)), 4), round(float(np.median(v)), 4), round(float(np.max(v)), 4)] for k, v in d.items()} for n_, d in res.items()}}))


if __name__ == "__main__":
    main()
