"""Plan images (include/ugnet.h, csrc/plan.cu): the compiled two-stage program as a relocatable blob that a host without
the Python lowering runs through the C ABI alone.  CPU: the image format and relocation coverage (stub engine).  GPU:
a plan exported from a PipelineRunner, loaded back through ug_plan_load, and run by a plain C program
(examples/run_plan.c, built with gcc) gives bit-identical masks / boxes / logits."""
import os
import struct
import subprocess
import sys

import numpy as np
import pytest
import torch

import ugnet_b200  # noqa: F401
from ugnet_b200 import engine as E
from ugnet_b200 import lower

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _parse(image):
    magic, n_allocs, n_ops, n_relocs, n_io, op_bytes = struct.unpack_from("<8sIIIIQ", image, 0)
    o = 32
    allocs = [struct.unpack_from("<QQQ", image, o + 24 * i) for i in range(n_allocs)]
    o += 24 * n_allocs + op_bytes * n_ops
    relocs = [struct.unpack_from("<IIIIQ", image, o + 24 * i) for i in range(n_relocs)]
    o += 24 * n_relocs
    ios = [struct.unpack_from("<32sIIQQ", image, o + 56 * i) for i in range(n_io)]
    return magic, allocs, n_ops, relocs, ios, op_bytes


def test_plan_image_format_and_relocation_coverage(monkeypatch):
    from test_lowering_cpu import _FakeEngine, _gnet_sd, _unet_sd
    import ctypes
    monkeypatch.setattr(E.Engine, "get", classmethod(lambda cls, device=0: _FakeEngine()))
    pipe = lower.PipelineRunner(_unet_sd(), _gnet_sd(), "cpu", micro_batch=1)
    image = pipe.export_plan(2)
    magic, allocs, n_ops, relocs, ios, op_bytes = _parse(image)
    ws = pipe.plan(2)
    assert magic == E.PLAN_MAGIC and op_bytes == ctypes.sizeof(E.Op) and n_ops == len(ws["program"].descs)
    # every non-null pointer field of every descriptor has exactly one relocation, inside its allocation
    want = sum(1 for d in ws["program"].descs for n, t in d._fields_ if t is E._vp and getattr(d, n))
    assert len(relocs) == want and want > 400
    for op, field, alloc, _, off in relocs:
        assert op < n_ops and field + 8 <= op_bytes and off <= allocs[alloc][0]
    names = sorted(n.split(b"\0")[0].decode() for n, *_ in ios)
    assert names == ["boxes", "cls_logits", "logits", "mask", "u8", "x_in"]
    sizes = {n.split(b"\0")[0].decode(): b for n, _, _, _, b in ios}
    assert sizes["x_in"] == 2 * 3 * 224 * 224 * 4 and sizes["mask"] == 2 * 224 * 224 and sizes["cls_logits"] == 2 * 6 * 4
    # the constants embedded in the image are exactly the two packed-weight blobs
    const = sorted(a[2] for a in allocs if a[2])
    assert const == sorted([pipe.unet.w_blob.numel(), pipe.gnet.w_blob.numel()])
    for nbytes, off, init in allocs:
        if init:
            assert off % 256 == 0 and off + init <= len(image)
    k = [a for a in allocs if a[2] == pipe.unet.w_blob.numel()][0]
    assert image[k[1]:k[1] + 4096] == pipe.unet.w_blob[:4096].numpy().tobytes()


@pytest.mark.gpu
def test_plan_runs_through_the_c_abi_and_a_plain_c_host(engine, tmp_path):
    from oracle import fixtures
    usd, gsd = fixtures.trained_unet_state(device="cuda"), fixtures.trained_googlenet_state(device="cuda")
    B = 4
    imgs, _, _ = fixtures.synth_images(B, seed=77)
    pipe = lower.PipelineRunner(usd, gsd, "cuda:0", micro_batch=2)
    masks, boxes, cls = pipe(torch.from_numpy(imgs).cuda())
    image = pipe.export_plan(B)
    # (1) through the ctypes binding of the C entry points, in this process
    plan = E.Plan(engine, image)
    assert set(plan.names) >= {"x_in", "mask", "boxes", "cls_logits"} and plan.device_bytes > 100e6
    plan.copy_in("x_in", torch.from_numpy(imgs))
    plan.run()
    m = plan.copy_out("mask", torch.empty((B, 224, 224), dtype=torch.uint8))
    b = plan.copy_out("boxes", torch.empty((B, 4), dtype=torch.int32))
    c = plan.copy_out("cls_logits", torch.empty((B, 6), dtype=torch.float32))
    assert torch.equal(m, masks.cpu()) and torch.equal(b, boxes.cpu()) and torch.equal(c, cls.cpu())
    with pytest.raises(RuntimeError):
        plan.copy_in("no_such_buffer", torch.zeros(4))
    plan.close()
    with pytest.raises(RuntimeError):
        E.Plan(engine, b"UGPLAN00" + image[8:])
    # (2) a plain C host: gcc + include/ugnet.h + libugnet.so, nothing else
    exe = tmp_path / "run_plan"
    libdir = os.path.join(ROOT, "unet-goolenet_b200")
    subprocess.run(["gcc", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "run_plan.c"), "-o",
                    str(exe), "-L", libdir, "-lugnet", f"-Wl,-rpath,{libdir}"], check=True)
    (tmp_path / "plan.bin").write_bytes(image)
    imgs.tofile(tmp_path / "images.f32")
    r = subprocess.run([str(exe), str(tmp_path / "plan.bin"), str(tmp_path / "images.f32"), str(B), str(tmp_path / "out")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert np.array_equal(np.fromfile(tmp_path / "out.mask.u8", np.uint8).reshape(B, 224, 224), masks.cpu().numpy())
    assert np.array_equal(np.fromfile(tmp_path / "out.boxes.i32", np.int32).reshape(B, 4), boxes.cpu().numpy())
    assert np.array_equal(np.fromfile(tmp_path / "out.cls.f32", np.float32).reshape(B, 6), cls.cpu().numpy())
    assert f"image {B - 1}: box" in r.stdout
