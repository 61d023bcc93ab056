"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement of the reference's `wavelet_enhance` (分类/test.py:17-63): grayscale image -> pseudo-RGB uint8
(R = normalised image, G = normalised Haar approximation, B = normalised Haar detail magnitude), the host
pre-processing step in front of the stage-2 path (test.py:127-130).

PARITY UNPINNED for this function: its wavelet arithmetic lives in PyWavelets (`pywt.wavedec2`, a dependency the
reference does not pin and that is absent from this image, so the reference function cannot be executed here) and
its resize in OpenCV (`cv2.resize`, present: 4.13).  The single-level Haar transform is restated from PyWavelets'
published algorithm (dwt2 = separable downsampling convolution, axis 0 then axis 1, mode 'symmetric', filters
dec_lo = [1/sqrt2, 1/sqrt2], dec_hi = [-1/sqrt2, 1/sqrt2], float32 arithmetic for float32 input,
out[o] = f[0]*x[2o+1] + f[1]*x[2o]); the resize calls cv2 itself.
"""
import numpy as np

F32 = np.float32
_C = F32(0.7071067811865476)


def _dwt_axis0(x):
    """Single-level Haar analysis along axis 0 in float32 ('symmetric' extension for an odd length)."""
    n = x.shape[0]
    if n % 2:
        x = np.concatenate([x, x[-1:]], axis=0)
    even, odd = x[0::2], x[1::2]
    lo = (_C * odd).astype(F32) + (_C * even).astype(F32)
    hi = ((-_C) * odd).astype(F32) + (_C * even).astype(F32)
    return lo.astype(F32), hi.astype(F32)


def haar_dwt2(x):
    """x: float32 [H, W] -> (cA, cH, cV, cD) as pywt.dwt2(x, 'haar') / pywt.wavedec2(x, 'haar', level=1)."""
    a, d = _dwt_axis0(x.astype(F32))
    aa, ad = (t.T for t in _dwt_axis0(a.T))
    da, dd = (t.T for t in _dwt_axis0(d.T))
    return aa, da, ad, dd


def _normalize(x):
    x = x - np.min(x)
    if np.max(x) != 0:
        x = x / np.max(x)
    return (x * 255).astype(np.uint8)


def wavelet_enhance(gray_img):
    """gray_img: [H, W] (or [1, H, W]) uint8 / float -> uint8 [3, H, W]; test.py:17-63 line by line."""
    import cv2
    if gray_img.ndim == 3:
        gray_img = gray_img[0]
    gray_img = gray_img.astype(F32)
    if gray_img.max() <= 1.0:
        gray_img = gray_img * F32(255.0)
    cA, cH, cV, cD = haar_dwt2(gray_img)
    high_freq = np.sqrt(cH ** 2 + cV ** 2 + cD ** 2)
    high_freq = cv2.resize(high_freq, gray_img.shape[::-1])
    low_freq = cv2.resize(cA, gray_img.shape[::-1])
    return np.stack([_normalize(gray_img), _normalize(low_freq), _normalize(high_freq)], axis=0)


def cv_resize_linear_f32(src, W, H):
    """NumPy restatement of cv2.resize(src, (W, H), INTER_LINEAR) for one float32 channel: horizontal pass, then
    vertical pass, each as the fused lerp  x0 + (x1 - x0) * f  (difference rounded to float32, then one FMA), the
    fraction f = (float)(fx - floor(fx)) with fx in double.  Bit-exact against cv2 4.13 (tests/test_oracle_wavelet.py);
    this is the form the CUDA kernel implements."""
    h, w = src.shape

    def coefs(n_in, n_out):
        scale = 1.0 / (n_out / n_in)
        idx = np.zeros(n_out, np.int64)
        a1 = np.zeros(n_out, F32)
        for d in range(n_out):
            fxd = (d + 0.5) * scale - 0.5            # double; OpenCV 4.x casts the FRACTION to float, not fx itself
            sx = int(np.floor(fxd))
            fx = F32(fxd - sx)
            if sx < 0:
                fx, sx = F32(0), 0
            if sx >= n_in - 1:
                fx, sx = F32(0), n_in - 1
            idx[d], a1[d] = sx, fx
        return idx, (F32(1.0) - a1).astype(F32), a1

    def lerp(x0, x1, f):   # fma(x1 - x0, f, x0): the product of two float32 is exact in float64
        diff = (x1 - x0).astype(F32).astype(np.float64)
        return (x0.astype(np.float64) + diff * f.astype(np.float64)).astype(F32)

    ix, _, ax1 = coefs(w, W)
    iy, _, ay1 = coefs(h, H)
    ix1, iy1 = np.minimum(ix + 1, w - 1), np.minimum(iy + 1, h - 1)
    rows = lerp(src[:, ix], src[:, ix1], ax1[None, :])
    return lerp(rows[iy, :], rows[iy1, :], ay1[:, None])
