"""GPU parity of the fused mesh regularisers (csrc/mesh_loss_api.cu): value and
gradient against the op-by-op PyTorch mirror (itself pinned to the reference's Python by
tests/test_host_golden.py) and against the reference-Python golden values directly."""
import os

import numpy as np
import pytest
import torch

import jrender_b200 as jr
from jrender_b200 import workloads as wl

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
VAL_TOL = 2e-5    # sums of 10^2..10^4 fp32 terms accumulated in a different order (warp tree + atomics)
GRAD_TOL = 5e-5   # of max |grad|: fp32 atomics vs the mirror's index_add order; forward-mode vs reverse-mode rounding


def _both(make_loss, v_np, dev):
    out = {}
    for fused in (True, False):
        loss_mod = make_loss().to(dev)
        loss_mod.fused = fused
        v = torch.from_numpy(v_np).to(dev).requires_grad_(True)
        val = loss_mod(v)
        w = torch.linspace(0.5, 1.5, val.numel(), device=dev).view_as(val)
        (val * w).sum().backward()
        out[fused] = (val.detach().cpu().numpy(), v.grad.cpu().numpy())
    return out


@pytest.mark.parametrize("which", ["laplacian", "flatten"])
def test_fused_losses_match_mirror_and_golden(which, cuda_device):
    g = np.load(os.path.join(G, "ref_host_loss_sphere280.npz"))
    f = torch.from_numpy(g["faces"])
    v0 = torch.from_numpy(g["vertices"][0])
    make = (lambda: jr.LaplacianLoss(v0, f)) if which == "laplacian" else (lambda: jr.FlattenLoss(f))
    out = _both(make, g["vertices"], cuda_device)
    ref = g[which]
    assert out[True][0].shape == ref.shape
    assert np.abs(out[True][0] - ref).max() <= VAL_TOL * np.abs(ref).max()          # vs the reference's own Python
    assert np.abs(out[True][0] - out[False][0]).max() <= VAL_TOL * np.abs(ref).max()
    assert np.abs(out[True][1] - out[False][1]).max() <= GRAD_TOL * np.abs(out[False][1]).max()


def test_fused_losses_large_mesh_average_and_graph(cuda_device):
    v, f = wl.sphere_by_faces(3280)
    rng = np.random.default_rng(9)
    vb = (v[None] + rng.normal(0, 0.01, (4,) + v.shape)).astype(np.float32)
    ft = torch.from_numpy(f)
    for make in (lambda: jr.LaplacianLoss(torch.from_numpy(v), ft, average=True), lambda: jr.FlattenLoss(ft, average=True)):
        out = _both(make, vb, cuda_device)
        assert abs(float(out[True][0]) - float(out[False][0])) <= VAL_TOL * abs(float(out[False][0]))
        assert np.abs(out[True][1] - out[False][1]).max() <= GRAD_TOL * np.abs(out[False][1]).max()
    # capture + replay: the fused ops enqueue on the capturing stream and never synchronise
    lap = jr.LaplacianLoss(torch.from_numpy(v), ft).to(cuda_device)
    flat = jr.FlattenLoss(ft).to(cuda_device)
    x = torch.from_numpy(vb).to(cuda_device).requires_grad_(True)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            x.grad = None
            (lap(x).mean() + flat(x).mean()).backward()
    torch.cuda.current_stream().wait_stream(s)
    eager = x.grad.clone()
    graph = torch.cuda.CUDAGraph()
    x.grad = None
    with torch.cuda.graph(graph):
        (lap(x).mean() + flat(x).mean()).backward()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.allclose(x.grad, eager, rtol=1e-4, atol=1e-6 * float(eager.abs().max()))
