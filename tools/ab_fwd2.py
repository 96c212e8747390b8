"""A/B of the forward kernels inside ONE library: variant 1 (lanes walk private lists) vs variant 2 (two-phase), outputs
compared bit for bit, kernel device times from the library's event profiler.
    [B200R_LIB=...] python tools/ab_fwd2.py [c3|c2|c5s]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ab_forward import kernel_times  # noqa: E402
from bench import WORKLOADS, build_inputs  # noqa: E402
from jrender_b200 import SoftRasterizeFunction, _lib  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
    kw = {}
    if wl == "c5s":   # demo2-like: silhouettes-ish params at 512^2 (sigma 1e-4, hard rgb), 16 views of the 3280-face sphere
        nf, H, bpg = 3280, 512, 16
        kw = dict(sigma_val=1e-4, aggr_func_rgb='hard')
        from jrender_b200 import workloads as wlm
        fv_h, tex_h = wlm.make_scene(nf, batch=bpg)
        grad_h = np.random.default_rng(2).uniform(-1, 1, (bpg, 4, H, H)).astype(np.float32)
    else:
        nf, H, bpg, _ = WORKLOADS[wl]
        fv_h, tex_h, grad_h = build_inputs(wl, 0, 1)
    L = _lib.lib()
    dev = torch.device("cuda:0")
    fv = torch.from_numpy(fv_h).to(dev).requires_grad_(True)
    tex = torch.from_numpy(tex_h).to(dev).requires_grad_(True)
    grad = torch.from_numpy(grad_h).to(dev)

    def step():
        fv.grad = None
        tex.grad = None
        SoftRasterizeFunction(image_size=H, **kw)(fv, tex).backward(grad)

    res = {"workload": wl, "lib": os.environ.get("B200R_LIB", "default")}
    outs = {}
    for variant in (1, 2):
        _lib.set_option("softras_fwd_variant", variant)
        fn = SoftRasterizeFunction(image_size=H, **kw)
        with torch.no_grad():
            sc, ag, ids = fn.raw(fv.detach(), tex.detach())
        outs[variant] = (sc.clone(), ag.clone(), ids.clone())
        for _ in range(3):
            step()
        res["variant%d" % variant] = kernel_times(L, 10, step)
    res["identical"] = {k: bool(torch.equal(a, b)) for k, a, b in zip(("soft_colors", "aggrs_info", "ids"), outs[1], outs[2])}
    if not all(res["identical"].values()):
        d = (outs[1][0] - outs[2][0]).abs()
        res["max_color_diff"] = float(d.max())
        res["ids_mismatch_frac"] = float((outs[1][2] != outs[2][2]).float().mean())
    _lib.set_option("softras_fwd_variant", 1)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
