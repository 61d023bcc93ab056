"""One pipeline program, a few passes, nothing else on the GPU between the passes — the ncu target of
scripts/gpu_evidence.sh.  Prints the number of engine kernels per pass so that the capture can skip the warm-up passes
exactly (`ncu -k <engine kernels> --launch-skip (passes-1)*L --launch-count L`).
usage: one_pass.py [--batch 128] [--micro-batch 128] [--passes 3] [--source-size 0] [--random-weights]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import ugnet_b200  # noqa: E402,F401
from ugnet_b200.lower import PipelineRunner  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--micro-batch", type=int, default=128)
ap.add_argument("--passes", type=int, default=3)
ap.add_argument("--source-size", type=int, default=0)
args = ap.parse_args()

from oracle import fixtures  # noqa: E402  (weights and images only: the trained-like fixture of the parity tests)
usd, gsd = fixtures.trained_unet_state(device="cuda"), fixtures.trained_googlenet_state(device="cuda")
imgs, _, _ = fixtures.synth_images(args.batch, seed=1234)
pipe = PipelineRunner(usd, gsd, "cuda:0", micro_batch=min(args.micro_batch, args.batch), cls_batch=args.batch)
S = args.source_size
ws = pipe.plan(args.batch, source=(S, S) if S else None)
x = torch.from_numpy(imgs).cuda()
if S:
    big = torch.nn.functional.interpolate(x, size=(S, S), mode="bilinear", align_corners=False)
    ws["src_u8"].copy_((big * 255).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1))
else:
    ws["x_in"].copy_(x)
torch.cuda.synchronize()
l0 = pipe.engine.launch_count
for _ in range(args.passes):
    ws["program"].run()
torch.cuda.synchronize()
per_pass = (pipe.engine.launch_count - l0) // args.passes
print(f"LAUNCHES_PER_PASS {per_pass} ops {ws['program'].num_launches} passes {args.passes} "
      f"mask_fg {ws['mask'].float().mean().item():.4f}")
