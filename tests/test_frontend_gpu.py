"""GPU parity of the device front-end (ug_resize_u8): bit-exact against Pillow itself — the
arithmetic the reference's CDDataAugmentation.transform runs (F.resize BILINEAR on a PIL image + to_tensor,
分类/util/data_utils.py:146-147) — for down-scaling (antialiased, up to 17 taps), up-scaling and identity."""
import numpy as np
import pytest
import torch
from PIL import Image

pytestmark = pytest.mark.gpu


def _pil(img, S):
    return np.asarray(Image.fromarray(img).resize((S, S), Image.BILINEAR))


@pytest.mark.parametrize("B,Hs,Ws,S", [(2, 512, 512, 224), (1, 300, 400, 224), (3, 224, 224, 224), (2, 100, 80, 224),
                                        (1, 1000, 700, 224), (1, 1792, 1792, 224), (2, 37, 511, 224), (1, 512, 512, 256),
                                        (1, 1, 1, 224)])
def test_resize_matches_pil_bit_exact(engine, B, Hs, Ws, S):
    from ugnet_b200.util.data_utils import resize_to_tensor
    rng = np.random.default_rng(Hs * 7 + Ws)
    src = rng.integers(0, 256, (B, Hs, Ws, 3), dtype=np.uint8)
    f32, u8 = resize_to_tensor(torch.from_numpy(src).cuda(), S, return_u8=True)
    torch.cuda.synchronize()
    for i in range(B):
        ref = _pil(src[i], S)
        assert np.array_equal(u8[i].cpu().numpy(), ref), f"image {i}: {(u8[i].cpu().numpy() != ref).sum()} pixels differ"
        ref_t = torch.from_numpy(ref).permute(2, 0, 1).float().div(255)      # F.to_tensor
        assert torch.equal(f32[i].cpu(), ref_t)


def test_resize_matches_oracle_and_smooth_images(engine):
    from oracle import roi_ref
    from ugnet_b200.util.data_utils import CDDataAugmentation
    yy, xx = np.mgrid[0:384, 0:640]
    img = np.stack([(xx * 255 // 639), (yy * 255 // 383), ((xx + yy) % 256)], -1).astype(np.uint8)
    got = CDDataAugmentation(img_size=224).transform(img)
    ref = roi_ref.pil_resize_bilinear_u8(img, 224)
    assert np.array_equal(np.round(got.cpu().numpy() * 255).astype(np.uint8).transpose(1, 2, 0), ref)


def test_resize_rejects_bad_args(engine):
    from ugnet_b200.util.data_utils import resize_to_tensor
    with pytest.raises(RuntimeError):
        resize_to_tensor(torch.zeros((1, 8, 8, 3), dtype=torch.uint8), 224)            # CPU tensor
    with pytest.raises(ValueError):
        resize_to_tensor(torch.zeros((1, 3, 8, 8), dtype=torch.uint8, device="cuda"), 224)
    with pytest.raises(RuntimeError):
        resize_to_tensor(torch.zeros((1, 2000, 16, 3), dtype=torch.uint8, device="cuda"), 224)   # > 8x


def test_pipeline_from_u8_sources(engine):
    """512x512 uint8 sources through the whole path == resizing with PIL on the host and feeding the float image."""
    from oracle import fixtures
    from ugnet_b200.lower import PipelineRunner
    usd = fixtures.procedural_state(fixtures.unet_template(), seed=7)
    gsd = fixtures.procedural_state(fixtures.googlenet_template(), seed=11)
    rng = np.random.default_rng(5)
    small = rng.integers(0, 256, (3, 32, 32, 3), dtype=np.uint8)
    src = np.stack([np.asarray(Image.fromarray(s).resize((512, 512), Image.BICUBIC)) for s in small])
    pipe = PipelineRunner(usd, gsd, "cuda:0", micro_batch=4)
    m1, b1, c1 = pipe(torch.from_numpy(src).cuda())
    host = np.stack([_pil(s, 224) for s in src]).transpose(0, 3, 1, 2).astype(np.float32) / np.float32(255)
    m2, b2, c2 = pipe(torch.from_numpy(host).cuda())
    assert torch.equal(m1, m2) and torch.equal(b1, b2) and torch.equal(c1, c2)
