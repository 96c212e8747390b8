"""Does the forward's speed depend on WHERE its workspace lies?  (Interleaved A/B runs showed identical code running
20-35 % apart on C2 / C5 depending on the library instance, i.e. on the workspace allocation.)
    python tools/ws_offset_probe.py c2|c5|c3"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ab_interleaved as ab  # noqa: E402
import math  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    fv_h, tex_h, H, sigma, rgb = ab.scene(name)
    dev = torch.device("cuda:0")
    B, nf = fv_h.shape[:2]
    K = 16
    L = ab.load(os.path.join(ROOT, "jrender_b200", "lib", "libb200raster.so"))
    fv, tex = torch.from_numpy(fv_h).to(dev), torch.from_numpy(tex_h).to(dev)
    out = torch.empty((B, 4, H, H), device=dev); aggr = torch.empty((B, 2, H, H), device=dev)
    ids = torch.empty((B, K, H, H), dtype=torch.int32, device=dev)
    nb = L.b200r_softras_workspace_bytes(B, nf, H)
    big = torch.empty(nb + (8 << 20), dtype=torch.uint8, device=dev)
    base = big.data_ptr()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    scal = (B, nf, 1, H, K, 1.0, 100.0, 1e-3, float(np.float32(sigma)), float(np.float32(1e-4)),
            float(np.float32(math.log(1.0 / 1e-4 - 1.0))), 2, rgb, 2, 0, 1)
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    res = {"workload": name, "ws_bytes": nb, "base_mod_2MiB": base % (2 << 20)}
    offs = [0, 256, 512, 1024, 4096, 65536, 1 << 20, (1 << 21) - (base % (1 << 21)), (1 << 21) - (base % (1 << 21)) + 256, 3 << 20, (3 << 20) + 768]
    for rnd in range(2):
        for off in offs:
            w = C.c_void_p(base + off)
            for _ in range(2):
                assert L.b200r_softras_forward(p(fv), p(tex), p(out), p(aggr), p(ids), None, w, nb, *scal, st) == 0
            L.b200r_profile_reset(); L.b200r_profile_enable(1)
            for _ in range(6):
                flush.fill_(1.0)
                L.b200r_softras_forward(p(fv), p(tex), p(out), p(aggr), p(ids), None, w, nb, *scal, st)
            torch.cuda.synchronize(); L.b200r_profile_enable(0)
            ms, cnt = C.c_double(0), C.c_longlong(0)
            L.b200r_profile_read(2, C.byref(ms), C.byref(cnt))
            res.setdefault(str(off), []).append(round(ms.value / max(1, cnt.value), 4))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
