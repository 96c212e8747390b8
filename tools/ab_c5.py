"""Kernel times of the C5 raster step (demo2: 3280-face sphere, 512^2, sigma 1e-4, silhouette mode) for the library
selected by B200R_LIB.    python tools/ab_c5.py [views]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ab_forward import kernel_times  # noqa: E402
from jrender_b200 import SoftRasterizeFunction, _lib, workloads as wl  # noqa: E402


def main():
    views = int(sys.argv[1]) if len(sys.argv) > 1 else 120
    L = _lib.lib()
    dev = torch.device("cuda:0")
    v, f = wl.sphere_by_faces(3280, radius=1.0)
    v = v * 0.5
    eyes = np.asarray([wl.get_points_from_angles(2.732 * 2, 30.0 * np.sin(b), 360.0 * b / views) for b in range(views)], np.float32)
    fv_h = wl.face_vertices(wl.perspective(wl.look_at(np.repeat(v[None], views, 0), eyes), 15.0), f)
    fv = torch.from_numpy(fv_h).to(dev).requires_grad_(True)
    tex = torch.ones((views, f.shape[0], 1, 3), device=dev, requires_grad=True)
    grad = torch.zeros((views, 4, 512, 512), device=dev)
    grad[:, 3] = torch.rand((views, 512, 512), device=dev) * 2 - 1

    def step():
        fv.grad = None
        tex.grad = None
        SoftRasterizeFunction(image_size=512, sigma_val=1e-4, aggr_func_rgb='none')(fv, tex).backward(grad)
    for _ in range(3):
        step()
    print(json.dumps({"views": views, **kernel_times(L, 10, step)}))


if __name__ == "__main__":
    main()
