"""Drop-in for the reference's networks/resnet.py (3-D ResNet encoder of CUNet / CTUNet).

Same constructors, forward signature and state_dict keys/shapes as the reference (resnet.py:82-245); the
torch layers below only HOLD the fp32 parameters (same construction order => same default init under the same
seed) — the arithmetic runs in the sm_100a kernels through hybrid_ctunet_b200.engine.Engine.
"""
from __future__ import annotations

from typing import Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from ._base import KernelModule, Program, from_cl, to_cl


def get_inplanes():
    return [32, 64, 128, 256]


def get_padding(kernel_size, stride):
    """resnet.py:52-64: (k - s + 1) / 2 per dim, truncated."""
    k, s = np.atleast_1d(kernel_size), np.atleast_1d(stride)
    p = (k - s + 1) / 2
    if np.min(p) < 0:
        raise AssertionError("padding value should not be negative, please change the kernel size and/or stride.")
    p = tuple(int(v) for v in p)
    return p if len(p) > 1 else p[0]


def get_output_padding(kernel_size, stride, padding):
    """resnet.py:66-80: 2p + s - k per dim."""
    k, s, p = np.atleast_1d(kernel_size), np.atleast_1d(stride), np.atleast_1d(padding)
    o = 2 * p + s - k
    if np.min(o) < 0:
        raise AssertionError("out_padding value should not be negative, please change the kernel size and/or stride.")
    o = tuple(int(v) for v in o)
    return o if len(o) > 1 else o[0]


class ConvHolder(nn.Sequential):
    """Parameter holder with MONAI's `Convolution(conv_only=True)` key layout: one child named `conv`."""

    def __init__(self, conv: nn.Module):
        super().__init__()
        self.add_module("conv", conv)


def get_conv_layer(spatial_dims: int, in_channels: int, out_channels: int, kernel_size=3, stride=1, act=None,
                   norm=None, dropout=None, groups: int = 1, bias: bool = False, conv_only: bool = True,
                   is_transposed: bool = False):
    """resnet.py:17-50 — returns the parameter holder for a bare Conv3d / ConvTranspose3d."""
    if spatial_dims != 3 or not conv_only or groups != 1:
        raise NotImplementedError("the CTUNet path uses 3-D, conv_only, groups=1 convolutions")
    padding = get_padding(kernel_size, stride)
    if is_transposed:
        conv = nn.ConvTranspose3d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                                  output_padding=get_output_padding(kernel_size, stride, padding), bias=bias)
    else:
        conv = nn.Conv3d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding, bias=bias)
    return ConvHolder(conv)


def _triple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v, v)


class Bottleneck(KernelModule):
    expansion = 4

    def __init__(self, in_planes: int, planes: int, spatial_dims: int = 3, stride=1, norm_name="INSTANCE",
                 dropout=None, downsample=None):
        super().__init__()
        if str(norm_name).upper() != "INSTANCE":
            raise NotImplementedError("the reference hard-wires InstanceNorm in the encoder (resnet.py:90,141,198)")
        self.conv1 = get_conv_layer(spatial_dims, in_planes, planes, kernel_size=1, stride=1)
        self.conv2 = get_conv_layer(spatial_dims, planes, planes, kernel_size=3, stride=stride)
        self.conv3 = get_conv_layer(spatial_dims, planes, planes * self.expansion, kernel_size=1, stride=1)
        self.downsample = downsample
        self.stride = stride

    def _program(self, eng, x):
        eng.stats.reset()
        a = to_cl(self._input(x))
        return Program([a], [eng.bottleneck("", a, _triple(self.stride), self.downsample is not None)])

    def forward(self, x):
        return self._call(x)[0]


class ResNet(KernelModule):
    def __init__(self, block, layers: Sequence[int], block_inplanes: Sequence[int], shortcut_type: str = "B",
                 n_input_channels: int = 1, conv1_t_size: int = 7,
                 DS_stride: tuple = ((2, 2, 1), (2, 2, 2), (2, 2, 2), (2, 2, 2)), no_max_pool: bool = True,
                 width_factor: float = 1.0, spatial_dims: int = 3, norm_name="INSTANCE"):
        super().__init__()
        if shortcut_type != "B" or not no_max_pool or n_input_channels != 1 or conv1_t_size != 7:
            raise NotImplementedError("CTUNet/CUNet build the encoder with shortcut B, no max-pool, 1 input channel, k7 stem")
        if tuple(map(tuple, DS_stride)) != ((2, 2, 1), (2, 2, 2), (2, 2, 2), (2, 2, 2)) or width_factor != 1.0:
            raise NotImplementedError("only the DS_stride / width the reference uses is implemented")
        block_inplanes = [int(x * width_factor) for x in block_inplanes]
        self.in_planes = 64
        self.no_max_pool = no_max_pool
        self.block_counts = list(layers)
        self.conv1 = get_conv_layer(spatial_dims, n_input_channels, self.in_planes, kernel_size=(7, 7, conv1_t_size),
                                    stride=DS_stride[0])
        self.layer1 = self._make_layer(block, block_inplanes[0], layers[0], shortcut_type)
        self.layer2 = self._make_layer(block, block_inplanes[1], layers[1], shortcut_type, stride=DS_stride[1])
        self.layer3 = self._make_layer(block, block_inplanes[2], layers[2], shortcut_type, stride=DS_stride[2])
        self.layer4 = self._make_layer(block, block_inplanes[3], layers[3], shortcut_type, stride=DS_stride[3])

    def _make_layer(self, block, planes, blocks, shortcut_type, stride=1):
        downsample = None
        if stride != 1 or self.in_planes != planes * block.expansion:
            # resnet.py:196-199: Sequential(conv1x1(stride), InstanceNorm) — the norm has no parameters
            downsample = nn.Sequential(get_conv_layer(3, self.in_planes, planes * block.expansion, kernel_size=1,
                                                      stride=stride), nn.Identity())
        mods = [block(in_planes=self.in_planes, planes=planes, stride=stride, downsample=downsample)]
        self.in_planes = planes * block.expansion
        for _ in range(1, blocks):
            mods.append(block(self.in_planes, planes))
        return nn.Sequential(*mods)

    def _program(self, eng, x):
        eng.stats.reset()
        return Program([None], eng.resnet("", self._input(x), self.block_counts))

    def forward(self, x):
        return list(self._call(x))


def generate_model(model_depth, **kwargs):
    """resnet.py:233-245."""
    assert model_depth in [50, 101, 152, 200]
    layers = {50: [3, 4, 6, 3], 101: [8, 9, 13, 3], 152: [8, 9, 30, 3], 200: [8, 25, 30, 3]}[model_depth]
    return ResNet(Bottleneck, layers, get_inplanes(), **kwargs)
