"""Does the GoogLeNet stage hide behind the next step's UNet kernels when it runs on a second stream?
Times (a) one stream: UNet(128) x2 then GoogLeNet(256)  (b) GoogLeNet(256) on a second stream concurrently with
UNet(128) x2 (what a software-pipelined serving loop would do: classify step i while segmenting step i+1)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ugnet_b200  # noqa
from oracle import fixtures
from ugnet_b200.lower import GoogLeNetRunner, UNetRunner

usd, gsd = fixtures.trained_unet_state(device="cuda"), fixtures.trained_googlenet_state(device="cuda")
u = UNetRunner(usd, "cuda:0", max_batch=128)
g = GoogLeNetRunner(gsd, "cuda:0", max_batch=256)
wu, wg = u.plan(128), g.plan(256, "u8")
imgs, _, _ = fixtures.synth_images(128, seed=3)
wu["x_in"].copy_(torch.from_numpy(imgs).cuda())
wg["in"].copy_((torch.from_numpy(imgs).cuda().repeat(2, 1, 1, 1) * 255).to(torch.uint8).permute(0, 2, 3, 1))
s2 = torch.cuda.Stream()
pu, pg = wu["program"], wg["program"]


def seq():
    pu.run(); pu.run(); pg.run()


def conc():
    ev = torch.cuda.Event()
    ev.record()
    with torch.cuda.stream(s2):
        s2.wait_event(ev)
        pg.run(stream=s2.cuda_stream)
    pu.run(); pu.run()
    torch.cuda.current_stream().wait_stream(s2)


def unet_only():
    pu.run(); pu.run()


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for rep in range(2):
    a, b, c = t(seq), t(conc), t(unet_only)
    print(f"one stream {a:.3f} ms | GoogLeNet on a second stream {b:.3f} ms | UNet x2 alone {c:.3f} ms  "
          f"-> hidden {100 * (a - b) / (a - c):.0f} % of the GoogLeNet stage", flush=True)
