"""A/B timing of kernel variants (device time per kernel from the library's event profiler) and,
when oracle/_ref is present, of the reference's own kernels on the same GPU.
    python tools/ab_forward.py [c3|c2]
"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, build_inputs  # noqa: E402
from jrender_b200 import SoftRasterizeFunction, _lib  # noqa: E402


def kernel_times(L, nsteps, step):
    L.b200r_profile_reset()
    L.b200r_profile_enable(1)
    for _ in range(nsteps):
        step()
    torch.cuda.synchronize()
    L.b200r_profile_enable(0)
    out = {}
    for kid, name in [(0, "setup"), (1, "coarse"), (2, "fwd"), (3, "bwd"), (4, "order"), (9, "bwd_finalize")]:
        ms, n = C.c_double(0), C.c_longlong(0)
        L.b200r_profile_read(kid, C.byref(ms), C.byref(n))
        out[name] = round(ms.value / max(1, n.value), 4)
    return out


def main():
    wlname = sys.argv[1] if len(sys.argv) > 1 else "c3"
    nf, H, bpg, desc = WORKLOADS[wlname]
    L = _lib.lib()
    dev = torch.device("cuda:0")
    fv_h, tex_h, grad_h = build_inputs(wlname, 0, 1)
    fv = torch.from_numpy(fv_h).to(dev).requires_grad_(True)
    tex = torch.from_numpy(tex_h).to(dev).requires_grad_(True)
    grad = torch.from_numpy(grad_h).to(dev)

    sigma = float(os.environ.get("AB_SIGMA", "1e-5"))   # 1e-4 = demo2's value, the stress point of SURVEY.md section 8d

    def step():
        fv.grad = None
        tex.grad = None
        SoftRasterizeFunction(image_size=H, sigma_val=sigma)(fv, tex).backward(grad)

    res = {"workload": desc, "sigma_val": sigma}
    for opt in os.environ.get("AB_OPTIONS", "").split(","):   # e.g. AB_OPTIONS=softras_heavy_faces=0
        if "=" in opt:
            k, v = opt.split("=")
            _lib.set_option(k, int(v))
            res[k] = int(v)
    for persistent in (1, 0):
        _lib.set_option("softras_fwd_persistent", persistent)
        for _ in range(3):
            step()
        res["persistent%d" % persistent] = kernel_times(L, 10, step)
    _lib.set_option("softras_fwd_persistent", 1)
    print(json.dumps(res), flush=True)

    from oracle import ref_gpu, softras as osr
    if ref_gpu.available() and not os.environ.get("AB_NO_REF"):
        P = osr.Params(image_size=H)
        fvd, texd = fv.detach(), tex.detach()

        def timed(fn, n=5):
            fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / n
        ref = {}
        fwd = ref_gpu.forward_t(fvd, texd, P)
        ids_t = fwd["faces_id_buffer"].permute(0, 2, 3, 1).contiguous()
        ref["naive_fwd_ms"] = timed(lambda: ref_gpu.forward_t(fvd, texd, P), 2)
        ref["c2f_bin64_fwd_ms"] = timed(lambda: ref_gpu.forward_t(fvd, texd, P, c2f_bin_size=64))
        ref["topk_bwd_ms"] = timed(lambda: ref_gpu.backward_t(fvd, texd, fwd, grad, P, ids_bhwk=ids_t))
        ref["transpose_ids_ms"] = timed(lambda: fwd["faces_id_buffer"].permute(0, 2, 3, 1).contiguous())
        print(json.dumps({"reference_kernels_same_gpu": ref, "batch": bpg}), flush=True)


if __name__ == "__main__":
    main()
