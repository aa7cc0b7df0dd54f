"""SparseTensor / PointTensor containers (torchsparse v2.0.0 `torchsparse/tensor.py` semantics)."""
import torch

__all__ = ["SparseTensor", "PointTensor"]


def _triple(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v, v)


class SparseTensor:
    def __init__(self, feats, coords, stride=1):
        self.feats = feats
        self.coords = coords
        self.stride = _triple(stride)
        self.cmaps = {}
        self.kmaps = {}

    # short aliases used all over the reference
    F = property(lambda s: s.feats, lambda s, v: setattr(s, "feats", v))
    C = property(lambda s: s.coords, lambda s, v: setattr(s, "coords", v))
    s = property(lambda s: s.stride)

    def _like(self, feats):
        out = SparseTensor(feats, self.coords, self.stride)
        out.cmaps, out.kmaps = self.cmaps, self.kmaps
        return out

    def __add__(self, other):
        return self._like(self.feats + other.feats)

    def cuda(self):
        return self

    def detach(self):
        self.feats = self.feats.detach()
        self.coords = self.coords.detach()
        return self


class PointTensor:
    def __init__(self, feats, coords, idx_query=None, weights=None):
        self.F = feats
        self.C = coords
        self.idx_query = idx_query if idx_query is not None else {}
        self.weights = weights if weights is not None else {}
        self.additional_features = {"idx_query": {}, "counts": {}}

    def cuda(self):
        return self

    def detach(self):
        self.F = self.F.detach()
        self.C = self.C.detach()
        return self
