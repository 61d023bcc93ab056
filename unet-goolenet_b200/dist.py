"""Multi-GPU plumbing of the path: contiguous batch shards per rank, one final gather.

Images are independent end to end, so each rank (one process per GPU, `torch.distributed`, NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests) runs the whole two-stage program on its slice; the only
exchange is the final gather of the per-rank masks (uint8), boxes and class logits (float32):

  gather_shards(local)            all-gather: every rank ends with the global tensor (what a caller that wants the result
                                  everywhere uses; `run_sharded`)
  gather_shards(local, dst=0)     gather to ONE rank (BASELINE.json: "a final NCCL gather of masks and logits"): only the
                                  destination receives, (world-1)/world of the all-gather's NVLink traffic disappears
  RootGather                      the serving-loop form: results of step i are staged into one of two slots and gathered
                                  to the root on a side stream while step i+1 computes (bench.py)
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous slice [lo, hi) of a global batch of n images owned by `rank` (n must divide evenly)."""
    if n % world:
        raise ValueError(f"global batch {n} is not divisible by world size {world}")
    per = n // world
    return rank * per, (rank + 1) * per


def gather_shards(local, out=None, group=None, dst=None):
    """Gather equal-sized per-rank tensors along dim 0, in rank order.  dst=None: all-gather (every rank returns the
    global tensor); dst=r: only rank r receives (returns the global tensor there, None elsewhere).  nccl and gloo."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    local = local.contiguous()
    need_out = dst is None or rank == dst
    if out is None and need_out:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if dst is None:
        if dist.get_backend(group) == "nccl":
            dist.all_gather_into_tensor(out, local, group=group)
        else:
            dist.all_gather(list(out.chunk(world, dim=0)), local, group=group)
        return out
    dist.gather(local, list(out.chunk(world, dim=0)) if rank == dst else None, dst=dst, group=group)
    return out if rank == dst else None


def run_sharded(pipeline_fn, images, group=None, dst=None):
    """Run `pipeline_fn(images_slice) -> (masks, boxes, logits)` on this rank's slice of the global batch and gather
    the global masks/boxes/logits (on every rank, or on rank `dst` only)."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_range(images.shape[0], rank, world)
    masks, boxes, logits = pipeline_fn(images[lo:hi])
    return tuple(gather_shards(t, group=group, dst=dst) for t in (masks, boxes, logits))


class RootGather:
    """Double-buffered gather-to-root that overlaps with the next step's kernels.

    submit(tensors) — on the caller's stream: copy each local tensor into staging slot s (device to device, so the
    program may overwrite its output buffers right away); on a side stream: wait for those copies, gather slot s to the
    root.  The caller's stream only ever waits when a slot is reused two submits later.  results() synchronizes the
    side stream and returns the root's global tensors of the LAST submit (None on the other ranks).
    On CPU tensors (gloo tests) everything is synchronous."""

    def __init__(self, like, world=None, rank=None, root=0, group=None, slots=2):
        self.group, self.root = group, root
        self.world = dist.get_world_size(group) if world is None else world
        self.rank = dist.get_rank(group) if rank is None else rank
        self.cuda = like[0].is_cuda
        self.slots = slots
        self.stage = [[torch.empty_like(t) for t in like] for _ in range(slots)]
        self.out = None
        if self.rank == root:
            self.out = [[torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
                         for t in like] for _ in range(slots)]
        self.n = 0
        if self.cuda:
            self.stream = torch.cuda.Stream(device=like[0].device)
            self.ready = [torch.cuda.Event() for _ in range(slots)]
            self.done = [torch.cuda.Event() for _ in range(slots)]

    def submit(self, tensors):
        s = self.n % self.slots
        if self.cuda:
            cur = torch.cuda.current_stream()
            if self.n >= self.slots:
                cur.wait_event(self.done[s])              # the gather that last read this slot has finished
            for dst, src in zip(self.stage[s], tensors):
                dst.copy_(src, non_blocking=True)
            self.ready[s].record(cur)
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(self.ready[s])
                for i, t in enumerate(self.stage[s]):
                    gather_shards(t, out=self.out[s][i] if self.out else None, group=self.group, dst=self.root)
                self.done[s].record(self.stream)
        else:
            for i, (dst, src) in enumerate(zip(self.stage[s], tensors)):
                dst.copy_(src)
                gather_shards(dst, out=self.out[s][i] if self.out else None, group=self.group, dst=self.root)
        self.n += 1

    def wait(self):
        """Make the caller's stream wait for every outstanding gather (end of a timed region)."""
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.stream)

    def results(self):
        self.wait()
        if self.cuda:
            torch.cuda.current_stream().synchronize()
        return self.out[(self.n - 1) % self.slots] if (self.out and self.n) else None
