"""Run one conv configuration three times (for ncu: -k regex:conv -s 2 -c 1).
usage: conv_one.py B,H,W,Cin,N,R variant [mode] [bn]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.argv, args = sys.argv[:1], sys.argv[1:]
import conv_prof as cp  # noqa: E402  (its sweep is skipped: SHAPES empty)
import torch  # noqa: E402
shape = tuple(int(v) for v in args[0].split(","))
cfg = dict(variant=int(args[1]))
if len(args) > 2:
    cfg["mode"] = int(args[2])
if len(args) > 3 and int(args[3]):
    cfg["bn"] = int(args[3])
if len(args) > 4:
    cfg["up"] = int(args[4])
d, keep = cp.make(*shape, **cfg)
for _ in range(3):
    cp.eng.run_op(d)
torch.cuda.synchronize()
print("ok", shape, cfg)
