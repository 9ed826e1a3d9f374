"""monai.networks.blocks.convolutions.Convolution, conv_only=True path only (MONAI 0.7.0 behaviour):
an nn.Sequential whose single child "conv" is a bare Conv3d / ConvTranspose3d."""
import torch.nn as nn


class Convolution(nn.Sequential):
    def __init__(self, spatial_dims, in_channels, out_channels, strides=1, kernel_size=3, adn_ordering="NDA",
                 act="PRELU", norm="INSTANCE", dropout=None, dropout_dim=1, dilation=1, groups=1, bias=True,
                 conv_only=False, is_transposed=False, padding=None, output_padding=None, dimensions=None):
        super().__init__()
        if spatial_dims != 3 or not conv_only:
            raise NotImplementedError("stub covers spatial_dims=3, conv_only=True (all the reference uses)")
        if padding is None:
            raise NotImplementedError("the reference always passes padding")
        if is_transposed:
            if output_padding is None:
                output_padding = 0
            conv = nn.ConvTranspose3d(in_channels, out_channels, kernel_size=kernel_size, stride=strides,
                                      padding=padding, output_padding=output_padding, groups=groups, bias=bias,
                                      dilation=dilation)
        else:
            conv = nn.Conv3d(in_channels, out_channels, kernel_size=kernel_size, stride=strides, padding=padding,
                             dilation=dilation, groups=groups, bias=bias)
        self.add_module("conv", conv)
