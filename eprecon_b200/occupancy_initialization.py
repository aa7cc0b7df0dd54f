"""Back_Project / Occupancy_Initialization — drop-ins for models/occupancy_initialization.py.

Back_Project.forward keeps the reference signature and 5-tuple (occupancy_initialization.py:189-261);
the whole body is three kernel launches (count, compact, gather) instead of ~25 ATen ops over
[V,C,N] intermediates.
"""
import torch
import torch.nn as nn

from . import ops


class Back_Project(nn.Module):
    """`materialize_grid=False` skips writing im_grid / mask (dead at the only call site,
    models/neucon_network.py:374) and returns None in their slots."""

    def __init__(self, dim, materialize_grid=True):
        super().__init__()
        self.materialize_grid = materialize_grid

    def forward(self, coords, origin, voxel_size, feats, KRcam, min_view_number):
        n_views, bs, c, h, w = feats.shape
        origin = origin.float().contiguous()
        KRcam = KRcam.float().contiguous()
        res = ops.backproject(coords.to(torch.int32).contiguous(), origin, voxel_size, ops.to_nhwc(feats.float()),
                              KRcam, min_view_number, mode="mean", want_src=True)
        if res is None:
            return None
        self.last = res  # survivors' source rows / visibility masks for fused callers
        im_grid = mask = None
        if self.materialize_grid:
            im_grid, mask = ops.backproject_grid(res, origin, voxel_size, KRcam, n_views, bs, h, w)
        return [res["feat"], res["coords"].to(coords.dtype), im_grid, mask, res["count"]]
