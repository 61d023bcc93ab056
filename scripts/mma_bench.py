import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("UG_DEV_LIB", "1")   # profiling hooks live in libugnet_dev.so (include/ugnet_dev.h)
import torch
import ugnet_b200  # noqa
from ugnet_b200 import engine as E
eng = E.Engine.get(0)
for N in (64, 128, 256):
    for n_acc in (1, 2, 4):
        if n_acc * N > 512: continue
        for cps in (1, 2):
            if cps * n_acc * N > 512: continue
            for distinct in (1,):
                out = (C.c_double * 2)()
                eng._check(eng.lib.ug_mma_microbench(eng.handle, N, n_acc, 2000, cps, distinct, out))
                mmas = 2000 * 4 * n_acc * cps
                print(f"N={N} n_acc={n_acc} ctas/SM={cps}: {out[0]:.1f} cycles/MMA per CTA (ideal {N/2:.0f}); "
                      f"launch {out[1]*1e3:.1f} us -> {out[1]*1e-3*1.965e9/mmas:.1f} cycles/MMA per SM @1.965GHz", flush=True)
