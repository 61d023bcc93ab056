"""Functional probe: does a TMA tensor map whose pixel stride (32 B) is smaller than its 64-element inner box (128 B)
deliver OVERLAPPING windows?  Runs the existing 1x1 implicit-GEMM kernels on an activation view with in_cstride = 16 and
Cin = 64: out[p, :] = W @ flat[p*16 : p*16 + 64] — a 4-tap x 16-channel row convolution if the windows overlap."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ugnet_b200  # noqa
from ugnet_b200 import engine as E, pack

eng = E.Engine.get(0)
g = torch.Generator(device="cuda").manual_seed(0)
P = 4096 * 8                                              # pixels (windows)
flat = torch.randn((P * 16 + 64,), generator=g, device="cuda").to(torch.bfloat16)
wt = torch.randn((64, 64), generator=g, device="cuda") * 0.125
wp = pack.pack_linear_weight(wt.cpu(), 64).cuda()
out = torch.zeros((P, 64), device="cuda", dtype=torch.bfloat16)
for variant in (1, 2, 5):
    d = E.ConvDesc()
    d.inp = flat.data_ptr(); d.in_cstride = 16; d.Cin = 64; d.B, d.H, d.W = 1, 1, P
    d.R = d.S = 1; d.pad = 0; d.w = wp.data_ptr(); d.N = 64; d.act = 0; d.mode = 0
    d.out = out.data_ptr(); d.out_cstride = 64; d.up = 1; d.BN = 64; d.variant = variant
    out.zero_()
    try:
        eng.run_op(d)
        torch.cuda.synchronize()
    except Exception as ex:
        print(f"variant {variant}: ERROR {ex}")
        continue
    win = flat[: P * 16 + 48].float().unfold(0, 64, 16)[:P]          # [P, 64] overlapping windows
    ref = win @ wt.to(torch.bfloat16).float().T
    err = (out.float() - ref).abs().max().item()
    print(f"variant {variant}: max err {err:.4f} (ref scale {ref.abs().max().item():.2f}) -> {'OVERLAPPING WINDOWS OK' if err < 0.1 else 'MISMATCH'}")
