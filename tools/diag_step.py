"""Diagnose step-time vs kernel-time gaps: per-step device times with / without L2 flush and
host enqueue time per step."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, build_inputs  # noqa: E402
from jrender_b200 import SoftRasterizeFunction  # noqa: E402

nf, H, bpg, desc = WORKLOADS["c3"]
dev = torch.device("cuda:0")
fv_h, tex_h, grad_h = build_inputs("c3", 0, 1)
fv = torch.from_numpy(fv_h).to(dev).requires_grad_(True)
tex = torch.from_numpy(tex_h).to(dev).requires_grad_(True)
grad = torch.from_numpy(grad_h).to(dev)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)


def step():
    fv.grad = None
    tex.grad = None
    SoftRasterizeFunction(image_size=H)(fv, tex).backward(grad)


for _ in range(5):
    step()
torch.cuda.synchronize()
for mode in ("noflush", "flush", "flush+sync_each"):
    ev, host = [], []
    for i in range(12):
        if mode != "noflush":
            flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        t0 = time.perf_counter()
        step()
        host.append((time.perf_counter() - t0) * 1e3)
        b.record()
        ev.append((a, b))
        if mode == "flush+sync_each":
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ev]
    print(mode, "device ms:", " ".join("%.2f" % x for x in ms), "| host enqueue ms:", " ".join("%.2f" % x for x in host), flush=True)
print("alloc stats: num_alloc_retries", torch.cuda.memory_stats()["num_alloc_retries"], "reserved MB", torch.cuda.memory_reserved() // 2**20)
