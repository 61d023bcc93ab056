"""CPU check of the next-round stem design: GoogLeNet conv1 (7x7, stride 2, pad 3, 3 channels) equals
a 4-tap (rows only) implicit GEMM with K = 64 per tap over OVERLAPPING 4-pixel windows of a space-to-depth image.

  s2d[y, x, (dy*2+dx)*3 + c] = xt[c, 2*y + dy, 2*x + dx]        12 channels, padded to 16; xt = transformed input
  out[n, y, x] = sum_{r=0..3} sum_{k=0..63} W2[n, r, k] * win[y + r - 2, x, k]
  win[yy, x, s*16 + ch] = s2d_padded[yy, x + s, ch]               s = 0..3: 4 adjacent pixels = 128 contiguous bytes
  (s2d_padded has 2 zero pixels on the left and 1 on the right; rows outside [0, 112) read as zero = TMA OOB fill)
  W2[n, r, s*16 + (dy*2+dx)*3 + c] = w[n, c, 2*(r-2) + dy + 3, 2*(s-2) + dx + 3]   (zero where the 7x7 index is out of range)
"""
import torch
import torch.nn.functional as F

torch.manual_seed(0)
B, S = 2, 224
w = torch.randn(64, 3, 7, 7, dtype=torch.float64)
x = torch.randn(B, 3, S, S, dtype=torch.float64)
ref = F.conv2d(x, w, stride=2, padding=3)                                   # [B, 64, 112, 112]

H2 = S // 2
s2d = torch.zeros(B, H2, H2 + 3, 16, dtype=torch.float64)                   # x-padding materialised: 2 left, 1 right
for dy in range(2):
    for dx in range(2):
        for c in range(3):
            s2d[:, :, 2:2 + H2, (dy * 2 + dx) * 3 + c] = x[:, c, dy::2, dx::2]

W2 = torch.zeros(64, 4, 64, dtype=torch.float64)
for r in range(4):
    for s in range(4):
        for dy in range(2):
            for dx in range(2):
                ky, kx = 2 * (r - 2) + dy + 3, 2 * (s - 2) + dx + 3
                if 0 <= ky < 7 and 0 <= kx < 7:
                    W2[:, r, s * 16 + (dy * 2 + dx) * 3:s * 16 + (dy * 2 + dx) * 3 + 3] = w[:, :, ky, kx]

# overlapping windows: win[b, yy, x, :] = s2d[b, yy, x:x+4, :].reshape(64) — as_strided = what the tensor map describes
win = s2d.as_strided((B, H2, H2, 64), (s2d.stride(0), s2d.stride(1), s2d.stride(2), 1))
winp = F.pad(win, (0, 0, 0, 0, 2, 1))                                       # rows -2 .. 112 (zero fill)
out = torch.zeros(B, H2, H2, 64, dtype=torch.float64)
for r in range(4):
    out += torch.einsum("byxk,nk->byxn", winp[:, r:r + H2], W2[:, r])
err = (out.permute(0, 3, 1, 2) - ref).abs().max().item()
print(f"max |s2d GEMM - conv2d| = {err:.3e}  (K = 4 taps x 64 = 256, {int((W2 != 0).sum() / 64)} of 256 columns non-zero)")
assert err < 1e-9
