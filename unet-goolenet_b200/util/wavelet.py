"""Device drop-in for the reference's `wavelet_enhance` (分类/test.py:17-63): grayscale image -> pseudo-RGB uint8
(R = normalised image, G = normalised Haar approximation, B = normalised Haar detail magnitude).

`wavelet_enhance(gray_img)` keeps the reference signature and return value ((3, H, W) uint8) for one image;
`wavelet_enhance_batch` is the batched form ([B,H,W] uint8 CUDA -> [B,H,W,3] uint8 CUDA, HWC as the reference
transposes it before `augm1.transform`, test.py:129-130), which feeds `util.data_utils.resize_to_tensor` and the
pipeline without leaving the GPU.  Kernels: csrc/wavelet.cu; there is no CPU path."""
import numpy as np
import torch

from .. import engine as E


@torch.no_grad()
def wavelet_enhance_batch(gray_u8):
    if gray_u8.device.type != "cuda":
        raise RuntimeError("ugnet: input must be a CUDA tensor (no CPU path)")
    if gray_u8.dtype != torch.uint8 or gray_u8.dim() != 3:
        raise ValueError(f"expected uint8 [B,H,W], got {gray_u8.dtype} {tuple(gray_u8.shape)}")
    g = gray_u8.contiguous()
    B, H, W = g.shape
    eng = E.Engine.get(g.device)
    nbytes = eng.lib.ug_wavelet_workspace_bytes(B, H, W)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=g.device)
    out = torch.empty((B, H, W, 3), dtype=torch.uint8, device=g.device)
    eng.run_op(E.WaveletDesc(g.data_ptr(), out.data_ptr(), ws.data_ptr(), nbytes, B, H, W))
    return out


def wavelet_enhance(gray_img, wavelet="haar", level=1, device="cuda"):
    """gray_img: (H, W) or (1, H, W) uint8 array -> (3, H, W) uint8 array, as the reference function."""
    if wavelet != "haar" or level != 1:
        raise NotImplementedError("the reference path calls wavelet_enhance(image) with haar, level 1 (test.py:128)")
    a = np.asarray(gray_img)
    if a.ndim == 3:
        a = a[0]
    if a.dtype != np.uint8:
        raise NotImplementedError("the reference path feeds cv2.imread(path, 0), i.e. uint8 (test.py:127)")
    out = wavelet_enhance_batch(torch.from_numpy(np.ascontiguousarray(a)).to(device)[None])[0]
    return out.permute(2, 0, 1).cpu().numpy()
