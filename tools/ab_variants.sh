#!/bin/bash
# A/B of development library variants (jrender_b200/build.py --variant): forward/backward kernel times at C3.
for lib in jrender_b200/lib/libb200raster*.so; do
  echo "== $lib"
  B200R_LIB=$PWD/$lib timeout 300 python - <<'PY'
import ctypes as C, sys, os, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
from bench import WORKLOADS, build_inputs
from jrender_b200 import SoftRasterizeFunction, _lib
from ab_forward import kernel_times
wl = os.environ.get("AB_WORKLOAD", "c3")
nf, H, bpg, desc = WORKLOADS[wl]
L = _lib.lib(); dev = torch.device("cuda:0")
fv_h, tex_h, grad_h = build_inputs(wl, 0, 1)
fv = torch.from_numpy(fv_h).to(dev).requires_grad_(True); tex = torch.from_numpy(tex_h).to(dev).requires_grad_(True)
grad = torch.from_numpy(grad_h).to(dev)
def step():
    fv.grad = None; tex.grad = None
    SoftRasterizeFunction(image_size=H)(fv, tex).backward(grad)
for _ in range(5): step()
print(kernel_times(L, 20, step))
PY
done
