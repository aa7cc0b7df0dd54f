"""back_project — drop-in for the reference's legacy operator ops/back_project.py:5-80.

Same arguments and return structure ([feat[N',C+1], coords[N',4] float32, count[N]] or None); the
projection, visibility, compaction and bilinear gather run in the sm_100a kernels of csrc/backproject.cu,
the extra normalised-depth channel (back_project.py:70-75) needs one global reduction over the
per-voxel mean depth the gather kernel emits.
"""
import torch

from . import ops


def back_project(coords, origin, voxel_size, feats, KRcam, min_view_number):
    n_views, bs, c, h, w = feats.shape
    res = ops.backproject(coords.to(torch.int32).contiguous(), origin.float().contiguous(), voxel_size,
                          ops.to_nhwc(feats.float()), KRcam.float().contiguous(), min_view_number, mode="mean",
                          want_zbar=True)
    if res is None:
        return None
    c_pad = (c + 1 + 3) // 4 * 4
    z = res["zbar"].unsqueeze(1)
    # per batch entry, as the reference normalises inside its batch loop (back_project.py:13,70-75)
    zn = torch.zeros_like(z)
    bidx = res["coords"][:, 0]
    for b in range(bs):
        sel = bidx == b if bs > 1 else slice(None)
        zb = z[sel]
        pos = zb[zb > 0]
        mu = pos.mean()
        s = torch.norm(pos - mu) + 1e-5
        znb = (zb - mu) / s
        znb[zb <= 0] = 0
        zn[sel] = znb
    del c_pad
    return [torch.cat([res["feat"], zn], dim=1), res["coords"].float(), res["count"]]
