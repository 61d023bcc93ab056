"""Poisoned guard bands around kernel outputs: compute-sanitizer is not available on the GPU pool, so the parity
tests carve every output out of a larger allocation and check the bytes on either side afterwards."""
import torch

GUARD = 4096   # elements on either side


def guarded(shape, fill, dtype, device="cuda"):
    """(view of `shape` filled with `fill`, checker) — call checker() after the kernel ran."""
    n = 1
    for v in shape:
        n *= v
    whole = torch.full((n + 2 * GUARD,), fill, device=device, dtype=dtype)

    def intact():
        torch.cuda.synchronize()
        assert bool((whole[:GUARD] == fill).all()), "kernel wrote before its output buffer"
        assert bool((whole[-GUARD:] == fill).all()), "kernel wrote past its output buffer"

    return whole[GUARD:GUARD + n].view(shape), intact
