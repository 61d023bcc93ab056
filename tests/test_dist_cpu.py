"""CPU, world_size 2, gloo: the shard/gather host logic of the multi-GPU path (the compute is stubbed by a
deterministic per-image function; the GPU pipeline itself is covered by the -m gpu tests)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ugnet_b200  # noqa: F401
from ugnet_b200.dist import RootGather, gather_shards, run_sharded, shard_range


def _fake_pipeline(x):
    masks = (x[:, 0] > 0.5).to(torch.uint8)
    boxes = torch.stack([x[:, 0].sum((1, 2)).int(), x[:, 1].sum((1, 2)).int(), x[:, 2].sum((1, 2)).int(),
                         torch.full((x.shape[0],), 7, dtype=torch.int32)], 1)
    logits = x.mean((2, 3)).repeat(1, 2)
    return masks, boxes, logits


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    imgs = torch.rand((8, 3, 16, 16), generator=g)
    masks, boxes, logits = run_sharded(_fake_pipeline, imgs)
    ref = _fake_pipeline(imgs)
    ok = torch.equal(masks, ref[0]) and torch.equal(boxes, ref[1]) and torch.equal(logits, ref[2])
    t = gather_shards(torch.full((3, 2), float(rank)))
    ok = ok and torch.equal(t, torch.tensor([0.0] * 6 + [1.0] * 6).reshape(6, 2))
    # gather to one rank only (the path's final collective as BASELINE.json words it)
    r0 = run_sharded(_fake_pipeline, imgs, dst=0)
    if rank == 0:
        ok = ok and all(torch.equal(a, b) for a, b in zip(r0, ref))
    else:
        ok = ok and all(t is None for t in r0)
    # double-buffered gather of a serving loop: three steps through two staging slots, last step's result on the root
    lo, hi = shard_range(8, rank, world)
    rg = None
    for step in range(3):
        m, b, lg = _fake_pipeline(imgs[lo:hi] * (0.5 + 0.25 * step))
        rg = rg or RootGather([m, b, lg], root=0)
        rg.submit([m, b, lg])
        m.zero_()                                        # the caller may overwrite its buffers right after submit
    res = rg.results()
    want = _fake_pipeline(imgs * 1.0)
    ok = ok and ((res is None) if rank else all(torch.equal(a, b) for a, b in zip(res, want)))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_shard_range():
    assert shard_range(512, 0, 8) == (0, 64) and shard_range(512, 7, 8) == (448, 512)
    with pytest.raises(ValueError):
        shard_range(10, 0, 4)


def test_sharded_gather_equals_unsharded():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
