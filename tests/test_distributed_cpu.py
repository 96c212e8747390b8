"""world_size-2 gloo test (CPU) of the multi-GPU host logic: batch sharding + the optional
all-gather of the image shards.  The per-rank "renderer" is the CPU oracle here (tests only)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.conftest import ROOT


def test_shard_range_partitions_the_batch():
    from jrender_b200.distributed import shard_range
    for B in (1, 4, 7, 32):
        for world in (1, 2, 3, 8):
            spans = [shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(32, 3, 8) == (12, 16)
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, batch, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from jrender_b200 import workloads as wl
    from jrender_b200.distributed import all_gather_images, shard_range
    from oracle import softras as osr
    fv, tex = wl.make_scene(280, batch=batch)
    lo, hi = shard_range(batch)
    P = osr.Params(image_size=32)
    img = torch.from_numpy(osr.forward(fv[lo:hi], tex[lo:hi], P, nthreads=1)["soft_colors"])
    full = all_gather_images(img, batch_size=batch)          # ragged when batch % world != 0
    full2 = all_gather_images(img)                           # sizes discovered by a collective
    assert torch.equal(full, full2)
    if batch % world == 0:   # the chunked / overlapped gather (synchronous on CPU) must give the same batch order
        from jrender_b200.distributed import OverlappedImageGather
        g = OverlappedImageGather(images_per_rank=hi - lo, image_shape=tuple(img.shape[1:]), device="cpu", chunk=1)
        for i in range(hi - lo):
            g.push(i, img[i:i + 1])
        assert torch.equal(g.result(), full)
        g = OverlappedImageGather(images_per_rank=hi - lo, image_shape=tuple(img.shape[1:]), device="cpu")   # one chunk
        g.push(0, img)
        assert torch.equal(g.result(), full)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), full.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [4, 3])
def test_two_rank_gloo_shard_and_gather(tmp_path, batch):
    port = 29500 + (os.getpid() % 1000) + batch
    mp.spawn(_worker, args=(2, port, batch, str(tmp_path)), nprocs=2, join=True)
    from jrender_b200 import workloads as wl
    from oracle import softras as osr
    fv, tex = wl.make_scene(280, batch=batch)
    ref = osr.forward(fv, tex, osr.Params(image_size=32), nthreads=1)["soft_colors"]
    got = np.load(os.path.join(tmp_path, "gathered.npy"))
    assert got.shape == ref.shape and np.array_equal(got, ref)
