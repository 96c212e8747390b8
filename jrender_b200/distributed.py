"""Multi-GPU plumbing: the rasterizer path shards over the batch dimension only.

One process per GPU (torchrun); rank r renders samples [lo, hi) of the global batch; there is
no data-path collective.  When a caller needs every image on every rank, `all_gather_images`
performs the single NCCL all-gather named by the north star (gloo on CPU for tests).
"""
import torch
import torch.distributed as dist


def shard_range(batch_size, rank=None, world_size=None):
    """Contiguous shard [lo, hi) of `batch_size` samples owned by `rank`; the first
    batch_size % world_size ranks get one extra sample."""
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    base, extra = divmod(batch_size, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_images(images, batch_size=None):
    """Gather per-rank image shards [b_r, ...] into the full batch [B, ...] on every rank.
    Equal shards use one all_gather_into_tensor; ragged shards are padded to the largest."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return images
    world = dist.get_world_size()
    if batch_size is None:
        n = torch.tensor([images.shape[0]], device=images.device)
        sizes = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(sizes, n)
        sizes = [int(s.item()) for s in sizes]
    else:
        sizes = [shard_range(batch_size, r, world)[1] - shard_range(batch_size, r, world)[0] for r in range(world)]
    images = images.contiguous()
    if len(set(sizes)) == 1:
        out = torch.empty((sizes[0] * world,) + tuple(images.shape[1:]), dtype=images.dtype, device=images.device)
        dist.all_gather_into_tensor(out, images)
        return out
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(images.shape[1:]), dtype=images.dtype, device=images.device)
    pad[:images.shape[0]] = images
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)
