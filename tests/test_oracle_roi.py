"""CPU: pin the ROI oracle (oracle/roi_ref.py) against PIL itself and against golden vectors produced by the
reference's own process_and_augment_roi (tests/golden/roi_golden.npz, see oracle/make_golden.py)."""
import os

import numpy as np
import pytest
from PIL import Image

from oracle import fixtures, roi_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden", "roi_golden.npz")


@pytest.mark.parametrize("h,w", [(3, 89), (110, 150), (224, 224), (30, 30), (224, 57), (61, 224), (1, 1)])
def test_resize_matches_pil_bit_exact(h, w):
    rng = np.random.default_rng(h * 1000 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((224, 224), Image.BILINEAR))
    assert np.array_equal(roi_ref.pil_resize_bilinear_u8(img, 224), ref)


@pytest.mark.parametrize("h,w", [(512, 512), (300, 400), (1000, 700), (225, 223), (1792, 448)])
def test_downscale_matches_pil_bit_exact(h, w):
    """The front-end case: Pillow's antialiased resample, support = in/out > 1."""
    rng = np.random.default_rng(h * 1000 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((224, 224), Image.BILINEAR))
    assert np.array_equal(roi_ref.pil_resize_bilinear_u8(img, 224), ref)


def test_bbox_edge_cases():
    H = W = 224
    m = np.zeros((H, W), np.uint8)
    assert roi_ref.bbox_from_mask(m) == (56, 56, 168, 168)                 # empty -> centred 112 square
    m[0, 0] = 1
    assert roi_ref.bbox_from_mask(m) == (0, 0, 30, 30)                     # 30 px right/below, exclusive end
    m[:] = 0; m[223, 223] = 1
    assert roi_ref.bbox_from_mask(m) == (193, 193, 224, 224)
    m[:] = 1
    assert roi_ref.bbox_from_mask(m) == (0, 0, 224, 224)
    m[:] = 0; m[100, 50] = 1; m[120, 90] = 1
    assert roi_ref.bbox_from_mask(m) == (20, 70, 120, 150)                 # asymmetric: -30 / +30 exclusive


def test_roi_matches_reference_golden():
    g = np.load(GOLD)
    masks, rois = g["masks"], g["rois"]
    imgs, _, _ = fixtures.synth_images(len(masks), seed=int(g["images_seed"]))
    imgs[:, 1] = np.clip(imgs[:, 1] * 0.8 + 0.1, 0, 1)
    imgs[:, 2] = np.clip(1.0 - imgs[:, 2], 0, 1)
    for i in range(len(masks)):
        roi, _ = roi_ref.roi_tensor(imgs[i], masks[i])
        got = np.round(roi * 255.0).astype(np.uint8)
        assert np.array_equal(got, rois[i]), f"case {i}"


@pytest.mark.skipif(not __import__("oracle.ref_import", fromlist=["x"]).available(),
                    reason="reference tree only exists in the build container")
def test_roi_oracle_matches_live_reference_on_random_masks():
    """The reference's own process_and_augment_roi (分类/util/roi.py:12-51, through a stub segmentation model that
    returns prescribed logits) against the restatement, on random blob / border / single-pixel / empty masks."""
    import torch
    from oracle import ref_import
    proc, Aug = ref_import.reference_roi()
    aug = Aug(img_size=224, ori_size=224, crop=None, p_hflip=0.0, p_vflip=0.0, color_jitter_params=None,
              long_mask=True)                                     # as 分类/test.py:113-116 constructs it

    class Stub(torch.nn.Module):
        def __init__(self, mask):
            super().__init__()
            self.mask = mask

        def forward(self, x):
            return (torch.from_numpy(self.mask).float() * 8.0 - 4.0)[None, None]

    rng = np.random.default_rng(123)
    masks = [np.zeros((224, 224), np.uint8)]
    for _ in range(9):
        m = np.zeros((224, 224), np.uint8)
        kind = rng.integers(0, 3)
        if kind == 0:                                             # a rectangle anywhere, possibly touching the border
            y0, x0 = rng.integers(0, 220, 2)
            m[y0:y0 + rng.integers(1, 120), x0:x0 + rng.integers(1, 120)] = 1
        elif kind == 1:                                           # a few isolated pixels
            for _ in range(rng.integers(1, 4)):
                m[rng.integers(0, 224), rng.integers(0, 224)] = 1
        else:                                                     # an ellipse
            yy, xx = np.mgrid[:224, :224]
            cy, cx, a, b = rng.integers(40, 184), rng.integers(40, 184), rng.integers(5, 80), rng.integers(5, 80)
            m[((yy - cy) / a) ** 2 + ((xx - cx) / b) ** 2 <= 1.0] = 1
        masks.append(m)
    imgs = rng.random((len(masks), 3, 224, 224), dtype=np.float32)
    for i, m in enumerate(masks):
        roi, se = proc(Stub(m), torch.from_numpy(imgs[i]), torch.device("cpu"), aug, f"{i}.png")
        want, box = roi_ref.roi_tensor(imgs[i], m)
        assert tuple(roi.shape) == (3, 224, 224)
        assert np.array_equal(np.round(roi.numpy() * 255).astype(np.uint8), np.round(want * 255).astype(np.uint8)), \
            f"case {i}, box {box}"
        assert np.abs(roi.numpy() - want).max() < 1e-6


@pytest.mark.skipif(not __import__("oracle.ref_import", fromlist=["x"]).available(),
                    reason="reference tree only exists in the build container")
@pytest.mark.parametrize("hs,ws", [(512, 512), (300, 400), (224, 224), (97, 131), (1000, 640)])
def test_frontend_oracle_matches_live_reference_transform(hs, ws):
    """the reference's CDDataAugmentation.transform in its inference configuration
    (分类/util/data_utils.py:92-148; test.py:113-116) on a uint8 HWC source of any size == the oracle's Pillow
    restatement followed by /255 (to_tensor), bit for bit."""
    from oracle import ref_import
    _, Aug = ref_import.reference_roi()
    aug = Aug(img_size=224, ori_size=224, crop=None, p_hflip=0.0, p_vflip=0.0, color_jitter_params=None,
              long_mask=True)
    rng = np.random.default_rng(hs * 7 + ws)
    src = rng.integers(0, 256, (hs, ws, 3), dtype=np.uint8)
    ref = aug.transform(src).numpy()                                # float32 [3,224,224]
    u8 = roi_ref.pil_resize_bilinear_u8(src, 224)                  # uint8 [224,224,3]
    want = np.transpose(u8, (2, 0, 1)).astype(np.float32) / np.float32(255)
    assert ref.shape == (3, 224, 224) and np.array_equal(ref, want)
