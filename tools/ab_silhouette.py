"""Kernel times of the silhouette-only path (colour aggregation compiled out) against the full render.
    python tools/ab_silhouette.py [c3|c2]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ab_forward import kernel_times  # noqa: E402
from bench import WORKLOADS, build_inputs  # noqa: E402
from jrender_b200 import SoftRasterizeFunction, _lib  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
    nf, H, bpg, desc = WORKLOADS[wl]
    L = _lib.lib()
    dev = torch.device("cuda:0")
    fv_h, tex_h, grad_h = build_inputs(wl, 0, 1)
    fv = torch.from_numpy(fv_h).to(dev).requires_grad_(True)
    tex = torch.from_numpy(tex_h).to(dev).requires_grad_(True)
    grad = torch.from_numpy(grad_h).to(dev)
    grad_a = grad.clone()
    grad_a[:, :3] = 0   # what autograd hands back for images[:, 3]
    res = {"workload": desc}
    for name, rgb, g in (("full_softmax", "softmax", grad), ("full_softmax_alpha_grad_only", "softmax", grad_a),
                         ("silhouette_none", "none", grad_a)):
        def step():
            fv.grad = None
            tex.grad = None
            SoftRasterizeFunction(image_size=H, aggr_func_rgb=rgb)(fv, tex).backward(g)
        for _ in range(3):
            step()
        res[name] = kernel_times(L, 10, step)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
