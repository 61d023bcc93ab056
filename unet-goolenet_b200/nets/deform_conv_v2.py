"""Parameter shell of the reference's DeformConv2d (分割/nets/deform_conv_v2.py:5-15).

CoordAtt3 instantiates it (basicUnet.py:213) but never calls it in forward, so the engine needs no kernel for
it; its five tensors only have to round-trip through load_state_dict(strict=True)."""
import torch
import torch.nn as nn


class DeformConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=1, dilation=1):
        super().__init__()
        k = kernel_size
        self.offset_conv = nn.Conv2d(in_channels, 2 * k * k, k, stride=stride, padding=padding, dilation=dilation)
        self.regular_conv = nn.Conv2d(in_channels, out_channels, k, stride=stride, padding=padding,
                                      dilation=dilation)
        self.bias = nn.Parameter(torch.zeros(out_channels))

    def forward(self, x):
        raise NotImplementedError("DeformConv2d is not on the inference hot path (never called by the reference "
                                  "forward); the engine only carries its parameters")
