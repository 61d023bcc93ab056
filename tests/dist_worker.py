"""Worker of tests/test_dist_gpu.py (one process per GPU under torch.distributed.run, NCCL): every rank runs the
two-stage pipeline on its contiguous slice of a seeded global batch and the results are gathered to rank 0 — through
gather_shards(dst=0), through the all-gather form, and through the double-buffered RootGather of the bench loop.
Rank 0 also runs the WHOLE batch on its own GPU and requires the gathered masks / boxes / logits to be bit-identical
("shard + gather equals the 1-GPU result bit for bit")."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ugnet_b200  # noqa: E402,F401
from oracle import fixtures  # noqa: E402
from ugnet_b200.dist import RootGather, gather_shards, shard_range  # noqa: E402
from ugnet_b200.lower import PipelineRunner  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 16 * world
    imgs, _, _ = fixtures.synth_images(n, seed=31)
    usd, gsd = fixtures.trained_unet_state(device=dev), fixtures.trained_googlenet_state(device=dev)
    pipe = PipelineRunner(usd, gsd, dev, micro_batch=8, cls_batch=16)
    lo, hi = shard_range(n, rank, world)
    m, b, c = pipe(torch.from_numpy(imgs[lo:hi]).to(dev))
    got_root = [gather_shards(t, dst=0) for t in (m, b, c)]
    got_all = [gather_shards(t) for t in (m, b, c)]
    rg = RootGather([m, b, c], root=0)
    for _ in range(3):                                   # slot reuse: three submits through two slots
        m2, b2, c2 = pipe(torch.from_numpy(imgs[lo:hi]).to(dev))
        rg.submit([m2, b2, c2])
    got_async = rg.results()
    ok = True
    if rank == 0:
        full = pipe(torch.from_numpy(imgs).to(dev))      # the 1-GPU result of the whole batch
        for name, got in (("gather", got_root), ("all_gather", got_all), ("root_gather_async", got_async)):
            for t, ref in zip(got, full):
                if not torch.equal(t, ref):
                    ok = False
                    print(f"MISMATCH {name}: {tuple(t.shape)} differs in {(t != ref).sum().item()} elements", flush=True)
        assert full[0].float().mean().item() > 0.01, "fixture masks should be non-trivial"
    else:
        ok = all(t is None for t in got_root) and got_async is None
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_OK" if flag.item() == 1 else "DIST_FAIL", f"world={world} images={n}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
