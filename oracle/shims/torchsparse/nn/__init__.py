"""Module layer of the torchsparse v2.0.0 shim: Conv3d / BatchNorm / ReLU."""
import math

import numpy as np
import torch
from torch import nn

from . import functional  # noqa: F401
from . import utils  # noqa: F401
from .functional import conv3d

__all__ = ["Conv3d", "BatchNorm", "ReLU"]


def _t(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v, v)


class Conv3d(nn.Module):
    """Parameter `kernel` [K,Cin,Cout] ([Cin,Cout] when K==1), no bias by default."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, dilation=1, bias=False,
                 transposed=False):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.dilation = _t(kernel_size), _t(stride), _t(dilation)
        self.transposed = transposed
        self.kernel_volume = int(np.prod(self.kernel_size))
        shape = (self.kernel_volume, in_channels, out_channels) if self.kernel_volume > 1 else (in_channels, out_channels)
        self.kernel = nn.Parameter(torch.zeros(*shape))
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)
        std = 1.0 / math.sqrt((out_channels if transposed else in_channels) * self.kernel_volume)
        self.kernel.data.uniform_(-std, std)
        if self.bias is not None:
            self.bias.data.uniform_(-std, std)

    def forward(self, x):
        return conv3d(x, self.kernel, self.bias, self.kernel_size, self.stride, self.dilation, self.transposed)


class BatchNorm(nn.BatchNorm1d):
    def forward(self, x):
        return x._like(super().forward(x.feats))


class ReLU(nn.ReLU):
    def forward(self, x):
        return x._like(super().forward(x.feats))
