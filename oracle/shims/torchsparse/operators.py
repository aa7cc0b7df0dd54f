import torch
from .tensor import SparseTensor


def cat(inputs):
    """Channel concat; coords/stride/maps of the first operand survive."""
    out = SparseTensor(torch.cat([t.feats for t in inputs], dim=1), inputs[0].coords, inputs[0].stride)
    out.cmaps, out.kmaps = inputs[0].cmaps, inputs[0].kmaps
    return out
