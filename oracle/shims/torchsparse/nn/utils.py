import numpy as np
import torch

__all__ = ["get_kernel_offsets"]


def _t(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v, v)


def get_kernel_offsets(size, stride=1, dilation=1, device="cpu"):
    """[K,3] int offsets.  Odd volume: z-outer/x-inner (MinkowskiEngine weight layout);
    even volume: x-outer/z-inner."""
    size, stride, dilation = _t(size), _t(stride), _t(dilation)
    axes = [np.arange(-size[k] // 2 + 1, size[k] // 2 + 1) * stride[k] * dilation[k] for k in range(3)]
    if int(np.prod(size)) % 2 == 1:
        offs = [[x, y, z] for z in axes[2] for y in axes[1] for x in axes[0]]
    else:
        offs = [[x, y, z] for x in axes[0] for y in axes[1] for z in axes[2]]
    return torch.tensor(np.asarray(offs, dtype=np.int32), dtype=torch.int, device=device)
