"""CPU: host-side logic of the drop-in shells — checkpoint formats of the reference (main.py:277, ROI_main.py:355:
`{'net': state_dict, 'optimizer': ..., 'epoch': ...}`), n-tile selection, batch chunking, writers."""
import io
import os

import numpy as np
import pytest
import torch

import ugnet_b200  # noqa: F401
from ugnet_b200 import pack


def _roundtrip(state):
    buf = io.BytesIO()
    torch.save(state, buf)
    buf.seek(0)
    return torch.load(buf, map_location="cpu")


def test_reference_checkpoint_format_loads_strict():
    """test.py:141-150 / predict.py:112-114: `model.load_state_dict(torch.load(path)['net'])` on all three shells."""
    from ugnet_b200.googlenet import GoogLeNetClassifier
    from ugnet_b200.nets.basicUnet_cls import UNetTaskAligWeight as ClsUNet
    from ugnet_b200.nets.basicUnet_new import UNetTaskAligWeight
    torch.manual_seed(0)
    for ctor, nkeys in ((lambda: UNetTaskAligWeight(n_channels=3, n_classes=1), 287),
                        (lambda: ClsUNet(n_channels=3, n_classes=1), 287),
                        (lambda: GoogLeNetClassifier(num_classes=6), 344)):
        src = ctor()
        ck = _roundtrip({"net": src.state_dict(), "optimizer": {"state": {}, "param_groups": []}, "epoch": 173})
        assert len(ck["net"]) == nkeys
        dst = ctor()
        res = dst.load_state_dict(ck["net"], strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        for k, v in src.state_dict().items():
            assert torch.equal(v, dst.state_dict()[k]), k
        # a checkpoint of the wrong network must fail loudly, as it does with the reference classes
        bad = dict(ck["net"])
        bad.pop(next(iter(bad)))
        with pytest.raises(RuntimeError):
            ctor().load_state_dict(bad, strict=True)


def test_shells_refuse_cpu_and_train_mode():
    from ugnet_b200.nets.basicUnet_cls import UNetTaskAligWeight as ClsUNet
    from ugnet_b200.nets.basicUnet_new import UNetTaskAligWeight
    for cls in (UNetTaskAligWeight, ClsUNet):
        m = cls(3, 1)
        with pytest.raises(RuntimeError):      # train mode (the engine is inference-only)
            m(torch.zeros(1, 3, 224, 224))
        m.eval()
        if not torch.cuda.is_available():
            with pytest.raises(RuntimeError):  # no CPU path
                m(torch.zeros(1, 3, 224, 224))


def test_choose_bn_covers_every_layer_width():
    widths = sorted({16, 24, 32, 48, 64, 96, 112, 128, 136, 144, 152, 160, 176, 192, 208, 224, 240, 256, 288, 320, 384,
                     512, 1024, 1536, 2048})
    for r in (1, 3):
        for n in widths:
            bn = pack.choose_bn(n, r=r)
            assert bn % 16 == 0 and 16 <= bn <= 256
            tiles = -(-n // bn)
            assert tiles * bn >= n
            if r == 3:
                # several n-tiles of the 3x3 multi-issuer kernel must be multiples of 64 wide (64-column store boxes)
                assert tiles == 1 or bn % 64 == 0, (n, bn)
            if r == 1 and tiles > 1:
                assert bn % 64 == 0 or n % bn == 0, (n, bn)
    # the GoogLeNet 1x1 widths that are not multiples of 128 get ONE fitted tile
    assert [pack.choose_bn(n) for n in (144, 160, 176, 192, 240)] == [144, 160, 176, 192, 240]
    assert pack.choose_bn(1536) == 256 and pack.choose_bn(512) == 128 and pack.choose_bn(96) == 96
    assert pack.choose_bn(2048, convt_cout=512) == 128 and pack.choose_bn(256, convt_cout=64) == 64


def test_pipeline_chunking():
    from ugnet_b200.lower import PipelineRunner

    class _Stub:
        micro_batch, cls_batch = 128, 256

    def chunks(n, mb=128, cb=256):
        s = _Stub()
        s.micro_batch, s.cls_batch = mb, cb
        return PipelineRunner._chunks(s, n)

    assert chunks(256) == [(0, 256)]
    assert chunks(2048) == [(i * 256, 256) for i in range(8)]
    assert chunks(300) == [(0, 256), (256, 44)]
    assert chunks(100) == [(0, 100)]
    assert chunks(129) == [(0, 128), (128, 1)]
    assert chunks(0) == []
    for n in (1, 7, 127, 128, 255, 257, 1000):   # a partition of [0, n) in order
        c = chunks(n, 4, 8)
        assert sum(k for _, k in c) == n and all(c[i][0] + c[i][1] == c[i + 1][0] for i in range(len(c) - 1))


def test_fold_bn_matches_batchnorm_eval():
    torch.manual_seed(1)
    conv = torch.nn.Conv2d(5, 7, 3, padding=1)
    bn = torch.nn.BatchNorm2d(7, eps=1e-3)
    bn.running_mean.normal_(); bn.running_var.uniform_(0.5, 2.0); bn.weight.data.normal_(); bn.bias.data.normal_()
    bn.eval()
    x = torch.randn(2, 5, 9, 9)
    with torch.no_grad():
        ref = bn(conv(x))
        scale, bias = pack.fold_bn(conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var, 1e-3)
        got = torch.nn.functional.conv2d(x, conv.weight, None, padding=1) * scale[None, :, None, None] + \
            bias[None, :, None, None]
    assert torch.allclose(got, ref, atol=1e-5, rtol=1e-5)


def test_conv1_space_to_depth_algebra():
    """GoogLeNet conv1 (7x7 stride 2 pad 3 over 3 channels) == the four-row-tap GEMM over overlapping 4-pixel windows of the
    2x2 space-to-depth image that the engine runs (ug_s2d_desc + pack.pack_conv1_s2d: K index r2*64 + s2*16 + (dy*2+dx)*3 + c).
    fp32 on the CPU; the bf16 rounding of the packed weight is the only difference allowed."""
    torch.manual_seed(4)
    S, cout = 20, 5
    x = torch.randn(2, 3, S, S)
    w = torch.randn(cout, 3, 7, 7) * 0.1
    ref = torch.nn.functional.conv2d(x, w.to(torch.bfloat16).float(), stride=2, padding=3)        # [2, cout, 10, 10]
    # packed image: pixel (Y, X) holds source pixels (2Y+dy-3, 2X+dx-3), channel (dy*2+dx)*3 + c, zero outside; Q = S/2 + 3
    Q = S // 2 + 3
    xp = torch.zeros(2, 3, 2 * Q, 2 * Q)
    xp[:, :, 3:3 + S, 3:3 + S] = x
    s2d = xp.reshape(2, 3, Q, 2, Q, 2).permute(0, 2, 4, 3, 5, 1).reshape(2, Q, Q, 12)              # [n, Y, X, (dy,dx,c)]
    s2d = torch.cat([s2d, torch.zeros(2, Q, Q, 4)], -1)                                             # 16 channels per pixel
    wp = pack.pack_conv1_s2d(w).float().reshape(cout, 4, 64)                                        # [co, r2, window of 4 px x 16 ch]
    O = S // 2
    out = torch.zeros(2, cout, O, O)
    for r2 in range(4):                       # row tap r2: A row of output (y, x) = the 64-element window at packed (y + r2, x .. x+3)
        win = torch.stack([s2d[:, r2:r2 + O, s2:s2 + O, :] for s2 in range(4)], 3).reshape(2, O, O, 64)
        out += torch.einsum("nyxk,ok->noyx", win, wp[:, r2])
    assert torch.allclose(out, ref, atol=1e-4, rtol=1e-4)


class _Loader(list):
    """A list of reference-style batches: {'image': float tensor, 'filename': [str]}."""


def test_result_txt_writer_matches_reference_format(tmp_path):
    """分类/test.py:74-96: '<name without .png> <class>' per line, sorted by the numeric file name."""
    from ugnet_b200.infer import inference_all_cls
    names = [["10.png", "2.png"], ["33.jpg", "1.png", "100.png"]]
    classes = {"10.png": 3, "2.png": 0, "33.jpg": 5, "1.png": 1, "100.png": 2}

    class Stub(torch.nn.Module):
        def forward(self, imgs):                     # the image's first pixel carries the class to predict
            return torch.nn.functional.one_hot(imgs[:, 0, 0, 0].long(), 6).float() * 4.0 - 1.0

    loader = _Loader()
    for batch in names:
        img = torch.zeros(len(batch), 3, 8, 8)
        for i, n in enumerate(batch):
            img[i, 0, 0, 0] = classes[n]
        loader.append({"image": img, "filename": batch})
    rec = inference_all_cls(Stub(), loader, "cpu", str(tmp_path))
    want = ["1 1", "2 0", "10 3", "33.jpg 5", "100 2"]
    assert rec == want
    assert (tmp_path / "result.txt").read_text() == "".join(w + "\n" for w in want)


def test_mask_png_writer_matches_reference_loop(tmp_path):
    """分割/predict.py:13-45: Segmentation_Results/<name minus .jpg>.png, red where mask == 1 on black."""
    from PIL import Image
    from ugnet_b200.infer import inference_all_seg
    rng = np.random.default_rng(0)
    masks = (rng.random((2, 224, 224)) > 0.7).astype(np.uint8)

    class Stub(torch.nn.Module):
        def forward_mask_boxes(self, imgs):
            return None, torch.from_numpy(masks), None

    loader = _Loader([{"image": torch.zeros(2, 3, 224, 224), "filename": ["7.jpg", "8.png"]}])
    out = inference_all_seg(Stub(), loader, "cpu", str(tmp_path))
    assert set(out) == {"7.jpg", "8.png"} and np.array_equal(out["7.jpg"], masks[0])
    for i, fn in enumerate(("7.png", "8.png.png")):          # the reference only strips '.jpg' (predict.py:32)
        got = np.asarray(Image.open(tmp_path / "Segmentation_Results" / fn))
        ref = Image.new("RGB", (224, 224), (0, 0, 0))        # the reference's own loop, restated
        for y in range(224):
            for x in range(224):
                if masks[i][y, x] == 1:
                    ref.putpixel((x, y), (255, 0, 0))
        assert got.shape == (224, 224, 3) and np.array_equal(got, np.asarray(ref))
