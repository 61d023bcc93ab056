"""CPU: properties of the wavelet oracle (oracle/wavelet_ref.py).  PyWavelets is absent from this image, so the
reference function cannot be run; what can be pinned is the Haar restatement against its closed form, perfect
reconstruction, the cv2 bilinear restatement against cv2 itself (bit-exact) and the output contract."""
import numpy as np

from oracle import wavelet_ref as w


def test_haar_closed_form_and_energy():
    rng = np.random.default_rng(0)
    x = rng.integers(0, 256, (37, 52)).astype(np.float32)
    cA, cH, cV, cD = w.haar_dwt2(x)
    assert cA.shape == (19, 26) and cH.shape == cV.shape == cD.shape == (19, 26)
    xe = np.concatenate([x, x[-1:]], 0).astype(np.float64)               # 'symmetric' extension of the odd axis
    a, b, c, d = xe[0::2, 0::2], xe[0::2, 1::2], xe[1::2, 0::2], xe[1::2, 1::2]
    assert np.allclose(cA, (a + b + c + d) / 2, atol=2e-4)
    assert np.allclose(np.abs(cH), np.abs(a + b - c - d) / 2, atol=2e-4)  # rows differ
    assert np.allclose(np.abs(cV), np.abs(a - b + c - d) / 2, atol=2e-4)  # columns differ
    assert np.allclose(np.abs(cD), np.abs(a - b - c + d) / 2, atol=2e-4)
    assert np.isclose((xe ** 2).sum(), (cA.astype(np.float64) ** 2 + cH.astype(np.float64) ** 2 +
                                        cV.astype(np.float64) ** 2 + cD.astype(np.float64) ** 2).sum(), rtol=1e-5)


def test_cv_bilinear_restatement_matches_cv2_bit_exact():
    import cv2
    rng = np.random.default_rng(1)
    for (h, ww, H, W) in [(112, 112, 224, 224), (151, 209, 301, 417), (256, 256, 512, 512), (2, 2, 4, 4), (3, 7, 5, 13)]:
        src = (rng.random((h, ww)) * 255).astype(np.float32)
        assert np.array_equal(cv2.resize(src, (W, H)), w.cv_resize_linear_f32(src, W, H))
    for (h, ww, H, W) in [(32, 1, 64, 2), (1, 5, 2, 9)]:     # one-pixel-wide sources: OpenCV takes another path, 1 ulp
        src = (rng.random((h, ww)) * 255).astype(np.float32)
        assert np.abs(cv2.resize(src, (W, H)) - w.cv_resize_linear_f32(src, W, H)).max() <= 2e-5


def test_output_contract():
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (64, 80), dtype=np.uint8)
    out = w.wavelet_enhance(img)
    assert out.shape == (3, 64, 80) and out.dtype == np.uint8
    assert out[0].min() == 0 and out[0].max() == 255                      # every channel is min-max normalised
    assert np.array_equal(out, w.wavelet_enhance(img[None]))              # (1, H, W) input form
    assert (w.wavelet_enhance(np.full((8, 8), 9, np.uint8)) == 0).all()   # flat image: all three channels zero
