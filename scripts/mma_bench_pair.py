"""tcgen05.mma.cta_group::2 issue-rate probe (csrc/microbench.cu, ug_mma_microbench_pair): CTA pairs, 1-4 issuing warps
in the leader CTA, N = 64 / 128 / 256.  Prints cycles per M=256 MMA per issuer and per pair; the per-SM floor of an
M=256 x N x 16 MMA is N/2 cycles."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("UG_DEV_LIB", "1")
import torch
import ugnet_b200  # noqa
from ugnet_b200 import engine as E
eng = E.Engine.get(0)
ITERS = 2000
for N in (64, 128, 256):
    for issuers in (1, 2, 3, 4):
        if issuers * N > 512:
            continue
        out = (C.c_double * 2)()
        try:
            eng._check(eng.lib.ug_mma_microbench_pair(eng.handle, N, issuers, ITERS, out))
        except Exception as ex:
            print(f"N={N} issuers={issuers}: ERR {ex}", flush=True)
            break
        per_pair = out[1] * 1e-3 * 1.965e9 / (ITERS * 4 * issuers)
        print(f"pair N={N} issuers={issuers}: {out[0]:.1f} cyc/MMA(M=256) per issuer; {per_pair:.1f} cyc/MMA per pair "
              f"@1.965GHz (floor {N // 2})", flush=True)
