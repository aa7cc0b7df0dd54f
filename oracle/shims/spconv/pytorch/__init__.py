"""SparseConvTensor + SubMConv3d / SparseConv3d restated for CPU.

SubMConv3d: output sites == input sites; out[j] = bias + sum_d W[:, d, :] . in[site_j + d - k//2]
(cross-correlation, weight layout [Cout, k, k, k, Cin], indices columns (b, i0, i1, i2))."""
import math

import numpy as np
import torch
from torch import nn

__all__ = ["SparseConvTensor", "SubMConv3d", "SparseConv3d"]


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size):
        self.features = features
        self.indices = indices
        self.spatial_shape = list(spatial_shape)
        self.batch_size = batch_size

    def replace_feature(self, feats):
        return SparseConvTensor(feats, self.indices, self.spatial_shape, self.batch_size)


def _t(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v, v)


def _linear_index(idx, shape):
    idx = idx.long()
    return ((idx[:, 0] * shape[0] + idx[:, 1]) * shape[1] + idx[:, 2]) * shape[2] + idx[:, 3]


class SubMConv3d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, indice_key=None, algo=None):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = _t(kernel_size)
        self.weight = nn.Parameter(torch.zeros(out_channels, *self.kernel_size, in_channels))
        if bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))
        else:
            self.register_parameter("bias", None)
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in = in_channels * int(np.prod(self.kernel_size))
            bound = 1.0 / math.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x):
        feats, idx, shape = x.features, x.indices.long(), x.spatial_shape
        n = feats.shape[0]
        lut = torch.full((x.batch_size * shape[0] * shape[1] * shape[2],), -1, dtype=torch.long)
        lut[_linear_index(idx, shape)] = torch.arange(n)
        out = torch.zeros(n, self.out_channels, dtype=feats.dtype)
        k0, k1, k2 = self.kernel_size
        for a in range(k0):
            for b in range(k1):
                for c in range(k2):
                    nb = idx.clone()
                    nb[:, 1] += a - k0 // 2
                    nb[:, 2] += b - k1 // 2
                    nb[:, 3] += c - k2 // 2
                    ok = ((nb[:, 1] >= 0) & (nb[:, 1] < shape[0]) & (nb[:, 2] >= 0) & (nb[:, 2] < shape[1])
                          & (nb[:, 3] >= 0) & (nb[:, 3] < shape[2]))
                    src = torch.full((n,), -1, dtype=torch.long)
                    src[ok] = lut[_linear_index(nb[ok], shape)]
                    j = torch.nonzero(src >= 0).squeeze(1)
                    if j.numel():
                        out[j] += feats[src[j]] @ self.weight[:, a, b, c, :].t()
        if self.bias is not None:
            out = out + self.bias
        return x.replace_feature(out)


class SparseConv3d(nn.Module):
    """Constructed by the reference (models/modules.py:227) but never run on the hot path."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, indice_key=None, algo=None):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(out_channels, *_t(kernel_size), in_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None

    def forward(self, x):
        raise NotImplementedError("SparseConv3d is not on the hot path")
