"""Minimal PointTensor: the attribute surface (`.F`, `.C`) of torchsparse.PointTensor that the reference's
callers construct (models/neucon_network.py:401, models/gru_fusion.py:341-346).  The drop-in modules accept any
object with these two attributes, including a real torchsparse.PointTensor."""


class PointTensor:
    def __init__(self, feats, coords, idx_query=None, weights=None):
        self.F = feats
        self.C = coords
        self.idx_query = idx_query if idx_query is not None else {}
        self.weights = weights if weights is not None else {}
        self.additional_features = {"idx_query": {}, "counts": {}}

    def cuda(self):
        self.F, self.C = self.F.cuda(), self.C.cuda()
        return self

    def detach(self):
        self.F, self.C = self.F.detach(), self.C.detach()
        return self
