"""tcgen05.mma issue-rate probe: several issuing warps in one CTA vs co-resident CTAs (see csrc/microbench.cu)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ugnet_b200  # noqa
from ugnet_b200 import engine as E
eng = E.Engine.get(0)
ITERS = 2000
for N in (64, 128, 256):
    for cps in (1, 2, 3, 4):
        for issuers in (1, 2, 4):
            for n_acc in (1, 2):
                if cps * issuers * n_acc * N > 512:
                    continue
                out = (C.c_double * 2)()
                eng._check(eng.lib.ug_mma_microbench2(eng.handle, N, n_acc, issuers, ITERS, cps, out))
                mmas = ITERS * 4 * n_acc * issuers * cps
                print(f"N={N} ctas/SM={cps} issuers={issuers} n_acc={n_acc}: {out[0]:.1f} cyc/MMA per issuer; "
                      f"{out[1]*1e-3*1.965e9/mmas:.1f} cyc/MMA per SM @1.965GHz (floor {N/2:.0f})", flush=True)
