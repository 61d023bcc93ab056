"""GPU parity of the device `wavelet_enhance` (ug_wavelet) against the oracle restatement of
分类/test.py:17-63 (oracle/wavelet_ref.py).  The oracle's wavelet step restates PyWavelets (absent here: parity of
that step is unpinned, see the oracle header) and its resize calls cv2 itself, which the kernel restates bit-exactly
(fused lerp, fraction cast from double).  Gate: all three channels bit-exact against the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _images(seed, B, H, W):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    out = []
    for b in range(B):
        base = 110 + 70 * np.sin(xx / (9.0 + b)) * np.cos(yy / (13.0 + b)) + rng.normal(0, 18, (H, W))
        out.append(np.clip(base, 0, 255).astype(np.uint8))
    return np.stack(out)


@pytest.mark.parametrize("B,H,W", [(3, 224, 224), (2, 512, 512), (2, 301, 417), (1, 64, 2), (1, 2, 9)])
def test_wavelet_matches_oracle(engine, B, H, W):
    from oracle import wavelet_ref
    from ugnet_b200.util.wavelet import wavelet_enhance_batch
    imgs = _images(H * 3 + W, B, H, W)
    got = wavelet_enhance_batch(torch.from_numpy(imgs).cuda()).cpu().numpy()
    for i in range(B):
        ref = wavelet_ref.wavelet_enhance(imgs[i]).transpose(1, 2, 0)
        diff = got[i].astype(np.int32) - ref.astype(np.int32)
        assert np.array_equal(got[i], ref), f"image {i}: {(diff != 0).sum()} of {diff.size} values differ, max {np.abs(diff).max()}"


def test_wavelet_special_cases(engine):
    from oracle import wavelet_ref
    from ugnet_b200.util.wavelet import wavelet_enhance
    flat = np.full((32, 48), 77, np.uint8)                      # max(x - min) == 0: normalize() keeps zeros
    assert np.array_equal(wavelet_enhance(flat), wavelet_ref.wavelet_enhance(flat))
    binary = (np.random.default_rng(1).random((40, 40)) > 0.5).astype(np.uint8)   # max <= 1: rescaled by 255
    got, ref = wavelet_enhance(binary), wavelet_ref.wavelet_enhance(binary)
    assert got.shape == (3, 40, 40) and np.array_equal(got, ref)
    with pytest.raises(NotImplementedError):
        wavelet_enhance(flat, wavelet="db2")


def test_wavelet_feeds_the_pipeline(engine):
    """test.py:127-131 on the device: gray -> wavelet_enhance -> resize 224 + to_tensor -> two-stage path."""
    from oracle import fixtures
    from ugnet_b200.lower import PipelineRunner
    from ugnet_b200.util.wavelet import wavelet_enhance_batch
    usd = fixtures.procedural_state(fixtures.unet_template(), seed=7)
    gsd = fixtures.procedural_state(fixtures.googlenet_template(), seed=11)
    pipe = PipelineRunner(usd, gsd, "cuda:0", micro_batch=2)
    rgb = wavelet_enhance_batch(torch.from_numpy(_images(3, 2, 300, 400)).cuda())
    masks, boxes, cls = pipe(rgb)
    assert masks.shape == (2, 224, 224) and boxes.shape == (2, 4) and cls.shape == (2, 6)
    assert torch.isfinite(cls).all()


def test_grade_images_end_to_end(engine, tmp_path):
    """infer.grade_images == composing the device stages by hand; records sorted by numeric file name."""
    from oracle import fixtures
    from ugnet_b200.infer import grade_images
    from ugnet_b200.lower import PipelineRunner
    from ugnet_b200.util.data_utils import resize_to_tensor
    from ugnet_b200.util.wavelet import wavelet_enhance_batch
    usd = fixtures.procedural_state(fixtures.unet_template(), seed=7)
    gsd = fixtures.procedural_state(fixtures.googlenet_template(), seed=11)
    pipe = PipelineRunner(usd, gsd, "cuda:0", micro_batch=4)
    gray = _images(11, 5, 256, 320)
    names = ["10.png", "9.png", "2.png", "33.png", "1.png"]
    recs = grade_images(pipe, gray, names, save_dir=str(tmp_path), batch_size=4)
    x = resize_to_tensor(wavelet_enhance_batch(torch.from_numpy(gray).cuda()), 224)
    ref = torch.argmax(pipe(x)[2], dim=1).cpu().numpy()
    want = sorted((f"{n.replace('.png', '')} {int(c)}" for n, c in zip(names, ref)), key=lambda r: int(r.split()[0]))
    assert recs == want
    assert (tmp_path / "result.txt").read_text().splitlines() == want
