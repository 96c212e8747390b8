"""End-to-end use of the drop-in API on the GPU: the reference's demo2 silhouette-fitting loop
(Renderer -> Lighting -> Transform -> SoftRasterizer, neg-IoU + Laplacian + flatten losses, Adam)
must actually optimise geometry through the hand-written backward."""
import pytest

pytestmark = pytest.mark.gpu


def test_demo2_deform_improves_iou(cuda_device):
    from examples.demo2_deform import run
    r = run(iters=100, image_size=64, batch_size=8, verbose=False)
    assert r["first_iou"] < 0.7, r
    assert r["final_iou"] > r["first_iou"] + 0.15, r          # measured: 0.55 -> 0.78 after 80, 0.88 after 200 iterations
    losses = [h[1] for h in r["history"]]
    assert losses[-1] < 0.7 * losses[0], r
