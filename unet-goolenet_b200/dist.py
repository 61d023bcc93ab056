"""Multi-GPU plumbing of the path: contiguous batch shards per rank, one final gather.

Images are independent end to end, so each rank (one process per GPU, `torch.distributed`, NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests) runs the whole two-stage program on its slice; the only
exchange is an all-gather of the per-rank masks (uint8) and class logits (float32) — equal-sized shards, so a
single `all_gather_into_tensor` per tensor."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous slice [lo, hi) of a global batch of n images owned by `rank` (n must divide evenly)."""
    if n % world:
        raise ValueError(f"global batch {n} is not divisible by world size {world}")
    per = n // world
    return rank * per, (rank + 1) * per


def gather_shards(local, out=None, group=None):
    """All-gather equal-sized per-rank tensors along dim 0, in rank order. Works with nccl and gloo."""
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    else:
        dist.all_gather(list(out.chunk(world, dim=0)), local.contiguous(), group=group)
    return out


def run_sharded(pipeline_fn, images, group=None):
    """Run `pipeline_fn(images_slice) -> (masks, boxes, logits)` on this rank's slice of the global batch and
    gather the global masks/boxes/logits on every rank."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(images.shape[0], rank, world)
    masks, boxes, logits = pipeline_fn(images[lo:hi])
    return gather_shards(masks, group=group), gather_shards(boxes, group=group), gather_shards(logits, group=group)
