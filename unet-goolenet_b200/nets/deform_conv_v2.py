"""Parameter container standing in for the reference's DeformConv2d (分割/nets/deform_conv_v2.py:5-15).

CoordAtt3 instantiates it (basicUnet.py:213) but never calls it in forward, so the engine needs no kernel for
it; its five tensors (`bias`, `offset_conv.{weight,bias}`, `regular_conv.{weight,bias}`) only have to round-trip
through load_state_dict(strict=True) with the reference's shapes."""
import torch
from torch import nn


class DeformConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=1, dilation=1):
        super().__init__()
        geometry = dict(kernel_size=kernel_size, stride=stride, padding=padding, dilation=dilation)
        # 2 offsets (dy, dx) per filter tap for the offset branch; the filter itself for the regular branch
        for attr, width in (("offset_conv", 2 * kernel_size ** 2), ("regular_conv", out_channels)):
            self.add_module(attr, nn.Conv2d(in_channels, width, **geometry))
        self.register_parameter("bias", nn.Parameter(torch.zeros(out_channels)))

    def forward(self, x):
        raise NotImplementedError("DeformConv2d is dead code in the reference forward (CoordAtt3 never calls it); "
                                  "the engine carries its parameters only")
