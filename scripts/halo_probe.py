"""Bring-up probe for the 3x3 halo kernel: which descriptor convention is correct, and how fast is it."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import ugnet_b200  # noqa
from ugnet_b200 import engine as E
from test_conv_gpu import run_conv
torch.backends.cudnn.allow_tf32 = False
eng = E.Engine.get(0)
CASES = [dict(B=2, H=16, W=16, Cin=64, N=64, R=3), dict(B=2, H=28, W=28, Cin=128, N=256, R=3),
         dict(B=3, H=14, W=14, Cin=512, N=512, R=3), dict(B=1, H=224, W=224, Cin=64, N=64, R=3),
         dict(B=2, H=56, W=56, Cin=256, N=128, R=3, mode=2), dict(B=2, H=32, W=48, Cin=64, N=64, R=3, mode=3),
         dict(B=2, H=14, W=14, Cin=24, N=64, R=3, in_extra=40, in_off=16),
         dict(B=3, H=14, W=14, Cin=256, N=208, R=3, out_extra=48, out_off=16),
         dict(B=2, H=112, W=112, Cin=128, N=64, R=3)]
for variant in (3, 4):
    for c in CASES:
        try:
            run_conv(eng, variant=variant, **c)
            print(f"variant {variant} {c}: OK", flush=True)
        except AssertionError as ex:
            print(f"variant {variant} {c}: MISMATCH {str(ex)[:150]}", flush=True)
        except Exception as ex:
            print(f"variant {variant} {c}: ERROR {str(ex)[:200]}", flush=True)
            if "CUDA" in str(ex) or "ugnet error -2" in str(ex):
                raise
