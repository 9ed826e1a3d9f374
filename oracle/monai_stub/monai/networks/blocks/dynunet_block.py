"""monai.networks.blocks.dynunet_block.UnetOutBlock (MONAI 0.7.0): 1x1x1 conv with bias, key `conv.conv.*`."""
import torch.nn as nn

from .convolutions import Convolution


class UnetOutBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, dropout=None):
        super().__init__()
        self.conv = Convolution(spatial_dims, in_channels, out_channels, strides=1, kernel_size=1, act=None,
                                norm=None, dropout=dropout, bias=True, conv_only=True, padding=0)

    def forward(self, inp):
        return self.conv(inp)
