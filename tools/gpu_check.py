"""Exploratory parity sweep (CUDA path vs CPU oracle) -- prints error statistics per config.
Used to SET the tolerances written in tests/; not a test itself.
    python tools/gpu_check.py [--quick]
"""
import itertools
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import softras as osr  # noqa: E402
from jrender_b200 import workloads as wl  # noqa: E402
from tests.util import run_cuda, run_oracle, sorted_ids  # noqa: E402


def compare(name, fv, tex, P, seed=2):
    H = P["image_size"]
    g = np.random.default_rng(seed).uniform(-1, 1, (fv.shape[0], 4, H, H)).astype(np.float32)
    t0 = time.time()
    ref = run_oracle(fv, tex, P, grad=g)
    t1 = time.time()
    got = run_cuda(fv, tex, P, grad=g)
    t2 = time.time()
    r = dict(name=name, cpu_s=round(t1 - t0, 3), gpu_s=round(t2 - t1, 3))
    r["ids_exact"] = bool(np.array_equal(ref["faces_id_buffer"], got["faces_id_buffer"]))
    r["ids_set_mismatch_px"] = int((sorted_ids(ref["faces_id_buffer"]) != sorted_ids(got["faces_id_buffer"])).any(1).sum())
    r["faces_info_exact"] = bool(np.array_equal(ref["faces_info"], got["faces_info"]))
    for k in ("soft_colors", "aggrs_info", "grad_faces", "grad_textures"):
        a, b = got[k].astype(np.float64), ref[k].astype(np.float64)
        d = np.abs(a - b)
        r[k] = dict(max_abs=float(d.max()), max_ref=float(np.abs(b).max()), nan=int(np.isnan(a).sum()),
                    n_gt_1e6=int((d > 1e-6 * max(1.0, np.abs(b).max())).sum()))
    r["alpha_max_abs"] = float(np.abs(got["soft_colors"][:, 3] - ref["soft_colors"][:, 3]).max())
    r["aggr1_exact"] = bool(np.array_equal(ref["aggrs_info"][:, 1], got["aggrs_info"][:, 1]))
    print(json.dumps(r), flush=True)
    return r


def main():
    quick = "--quick" in sys.argv
    res = []
    fv, tex = wl.make_scene(280, batch=2)
    res.append(compare("sphere280_default_64", fv, tex, osr.Params(image_size=64)))
    res.append(compare("sphere280_default_250", fv, tex, osr.Params(image_size=250)))
    res.append(compare("sphere280_sigma1e-4_128", fv, tex, osr.Params(image_size=128, sigma_val=1e-4)))
    for dist, rgb, alpha in itertools.product(["hard", "barycentric", "euclidean"], ["hard", "softmax", "none"], ["hard", "sum", "prod"]):
        P = osr.Params(image_size=96, dist_func=dist, aggr_func_rgb=rgb, aggr_func_alpha=alpha, sigma_val=3e-5)
        res.append(compare("modes_%s_%s_%s" % (dist, rgb, alpha), fv, tex, P))
    fvv, texv = wl.make_scene(280, batch=1, texture_type="vertex")
    for rgb in ["hard", "softmax"]:
        res.append(compare("vertex_%s" % rgb, fvv, texv, osr.Params(image_size=96, texture_type="vertex", aggr_func_rgb=rgb)))
    fv5, tex5 = wl.make_scene(280, batch=1, texture_res=5)
    for rgb in ["hard", "softmax"]:
        res.append(compare("surfaceT25_%s" % rgb, fv5, tex5, osr.Params(image_size=96, aggr_func_rgb=rgb)))
    fvr, texr = wl.random_triangles(2, 300, seed=3)
    res.append(compare("random300_K4", fvr, texr, osr.Params(image_size=80, max_faces_per_pixel_for_grad=4, sigma_val=1e-4)))
    res.append(compare("random300_K64", fvr, texr, osr.Params(image_size=80, max_faces_per_pixel_for_grad=64, sigma_val=1e-4)))
    res.append(compare("random300_nofillback", fvr, texr, osr.Params(image_size=80, fill_back=False)))
    if not quick:
        fv3, tex3 = wl.make_scene(3280, batch=1)
        res.append(compare("sphere3280_256", fv3, tex3, osr.Params(image_size=256)))
        fv4, tex4 = wl.make_scene(39200, batch=1)
        res.append(compare("sphere39200_256", fv4, tex4, osr.Params(image_size=256)))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/gpu_check.json", "w"), indent=1)


if __name__ == "__main__":
    main()
