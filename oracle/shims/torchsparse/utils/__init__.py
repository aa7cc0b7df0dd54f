"""`from torchsparse.utils import *` in the reference (models/modules.py:6) only needs to succeed."""
__all__ = ["make_ntuple"]


def make_ntuple(x, ndim=3):
    return tuple(x) if isinstance(x, (tuple, list)) else (x,) * ndim
