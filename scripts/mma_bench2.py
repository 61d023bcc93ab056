"""tcgen05.mma issue-rate probe: several issuing warps in one CTA vs co-resident CTAs, plain vs halo-layout A
descriptors, and the TMEM column distance between the accumulators of concurrent chains (csrc/microbench.cu)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("UG_DEV_LIB", "1")   # profiling hooks live in libugnet_dev.so (include/ugnet_dev.h)
import torch
import ugnet_b200  # noqa
from ugnet_b200 import engine as E
eng = E.Engine.get(0)
ITERS = 2000


def run(N, cps, issuers, n_acc=1, a_off=0, a_sbo=1024, acc_stride=0):
    out = (C.c_double * 2)()
    eng._check(eng.lib.ug_mma_microbench2(eng.handle, N, n_acc, issuers, ITERS, cps, a_off, a_sbo, acc_stride, out))
    mmas = ITERS * 4 * n_acc * issuers * cps
    print(f"N={N} ctas/SM={cps} issuers={issuers} n_acc={n_acc} a_off={a_off} sbo={a_sbo} acc_stride={acc_stride}: "
          f"{out[0]:.1f} cyc/MMA per issuer; {out[1]*1e-3*1.965e9/mmas:.1f} cyc/MMA per SM @1.965GHz "
          f"(floor {N/2:.0f})", flush=True)


for N in (64, 128):
    for issuers in (2, 4):
        for stride in (0, 64, 96, 128, 192, 256, 320, 384):
            if stride and stride < N:
                continue
            if (issuers - 1) * (stride or N) + N > 512:
                continue
            run(N, 1, issuers, acc_stride=stride)
run(32, 1, 2, acc_stride=0); run(32, 1, 2, acc_stride=256); run(256, 1, 2, acc_stride=0)
