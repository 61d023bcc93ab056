"""Time ug_chanstats on the four CoordAtt3 maps at B=128: python scripts/chanstats_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ugnet_b200  # noqa
from ugnet_b200 import engine as E

eng = E.Engine.get(0)
B, S = 128, 16
tot = 0.0
for size, C in ((28, 256), (56, 128), (112, 64), (224, 64)):
    x = torch.randn((B, size * size, C), device="cuda").to(torch.bfloat16)
    ps, pm = torch.empty((B, S, C), device="cuda"), torch.empty((B, S, C), device="cuda")
    d = E.ChanStatsDesc(x.data_ptr(), C, C, B, size * size, S, ps.data_ptr(), pm.data_ptr())
    for _ in range(3):
        eng.run_op(d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.run_op(d)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    tot += ms
    print(f"chanstats {size}x{size}x{C}: {ms:.4f} ms  {x.numel() * 2 / ms / 1e6:.0f} GB/s")
print(f"total {tot:.4f} ms")
