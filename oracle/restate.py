"""TEST INFRASTRUCTURE ONLY — self-contained CPU restatement of the EPRecon feature-volume hot path.

Functional style (explicit state dicts, no nn.Module mirror) so that it can travel to the GPU box,
where /root/reference does not exist.  Each function cites the reference lines it follows; the
restatement is pinned against the UNMODIFIED reference (run over oracle/shims in the build container)
by the fixtures in tests/golden/ (made by tests/golden/make_golden.py).  The third-party sparse-conv
semantics underneath (torchsparse v2.0.0 / spconv) are restated from their published algorithms and
are PARITY UNPINNED — see oracle/shims/torchsparse/__init__.py.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
if _SHIMS not in sys.path:
    sys.path.insert(0, _SHIMS)

f32 = np.float32


# ================================================================================ back-projection
def project_views(coords, origin, voxel_size, krcam, H, W):
    """Visibility arithmetic of models/occupancy_initialization.py:214-228 (same maths at :88-102 and
    ops/back_project.py:22-36), in numpy fp32 with one rounding per op and the 4-term dot product summed
    left to right (the order the CUDA kernel fixes; the reference leaves it to sgemm).

    coords int [N,4] (b,x,y,z); origin [bs,3]; krcam [V,bs,4,4].  Returns gx, gy, z [V,N] fp32, vis [V,N] bool.
    """
    c = coords.detach().cpu().numpy()
    org = origin.detach().cpu().numpy().astype(f32)
    kr = krcam.detach().cpu().numpy().astype(f32)
    b = c[:, 0].astype(np.int64)
    vs = f32(voxel_size)
    with np.errstate(all="ignore"):
        wx = c[:, 1].astype(f32) * vs + org[b, 0]
        wy = c[:, 2].astype(f32) * vs + org[b, 1]
        wz = c[:, 3].astype(f32) * vs + org[b, 2]
        P = kr[:, b]  # [V,N,4,4]
        X = ((P[:, :, 0, 0] * wx + P[:, :, 0, 1] * wy) + P[:, :, 0, 2] * wz) + P[:, :, 0, 3]
        Y = ((P[:, :, 1, 0] * wx + P[:, :, 1, 1] * wy) + P[:, :, 1, 2] * wz) + P[:, :, 1, 3]
        Z = ((P[:, :, 2, 0] * wx + P[:, :, 2, 1] * wy) + P[:, :, 2, 2] * wz) + P[:, :, 2, 3]
        u = X / Z
        v = Y / Z
        gx = (f32(2) * u) / f32(W - 1) - f32(1)
        gy = (f32(2) * v) / f32(H - 1) - f32(1)
        vis = (np.abs(gx) <= 1) & (np.abs(gy) <= 1) & (Z > 0)
    return gx.astype(f32), gy.astype(f32), Z.astype(f32), vis


def backproject(coords, origin, voxel_size, feats, krcam, min_view_number, mode="mean", min_valid=1):
    """Back_Project.forward (models/occupancy_initialization.py:189-261) / init-stage variance
    (:79-128) / legacy depth channel (ops/back_project.py:57-75) in one function.

    feats [V,bs,C,H,W].  Returns None on a degenerate fragment, else dict(feat, coords, count, im_grid, mask,
    zbar).  Sampling uses ATen grid_sample itself (the reference's op); masked mean / population
    variance follow the reference's op order.
    """
    V, bs, C, H, W = feats.shape
    gx, gy, Z, vis = project_views(coords, origin, voxel_size, krcam, H, W)
    count = vis.sum(0).astype(f32)
    keep = count >= min_view_number
    bidx = coords[:, 0].cpu().numpy()
    for b in range(bs):
        if int((keep & (bidx == b)).sum()) < min_valid:
            return None
    kidx = np.nonzero(keep)[0]
    grid = torch.from_numpy(np.stack([gx[:, kidx], gy[:, kidx]], -1))  # [V,M,2]
    mask = torch.from_numpy(vis[:, kidx])
    zk = torch.from_numpy(Z[:, kidx]).clone()
    kb = torch.from_numpy(bidx[kidx]).long()
    M = kidx.shape[0]
    sampled = torch.zeros(V, C, M)
    for b in range(bs):
        sel = torch.nonzero(kb == b).squeeze(1)
        if sel.numel() == 0:
            continue
        s = F.grid_sample(feats[:, b].float(), grid[:, sel].view(V, 1, -1, 2), padding_mode="zeros",
                          align_corners=True).view(V, C, -1)
        sampled[:, :, sel] = s
    sampled[mask.unsqueeze(1).expand(-1, C, -1) == False] = 0  # noqa: E712
    zk[mask == False] = 0  # noqa: E712
    cnt = mask.sum(0)
    if mode == "mean":
        denom = cnt.clone()
        denom[denom == 0] = 1
        feat = (sampled.sum(0) / denom.unsqueeze(0)).permute(1, 0).contiguous()
    else:  # masked mean + population variance over views (occupancy_initialization.py:124-128)
        fm = sampled.permute(2, 0, 1)  # [M,V,C]
        mk = mask.transpose(0, 1)
        mean = (mk.unsqueeze(2) * fm).sum(1) / mk.sum(1).unsqueeze(1)
        feat = (mk.unsqueeze(2) * ((fm - mean.unsqueeze(1)) ** 2)).sum(1) / mk.sum(1).unsqueeze(1)
    denom = cnt.clone().float()
    denom[denom == 0] = 1
    zbar = zk.sum(0) / denom
    return {"feat": feat, "coords": coords[torch.from_numpy(kidx)], "count": torch.from_numpy(count),
            "im_grid": grid, "mask": mask, "zbar": zbar, "src": torch.from_numpy(kidx)}


def legacy_depth_channel(zbar):
    """ops/back_project.py:70-75: normalise mean depth by the global mean and an L2-norm 'std'."""
    z = zbar.view(-1, 1)
    pos = z[z > 0]
    mu = pos.mean()
    s = torch.norm(pos - mu) + 1e-5
    zn = (z - mu) / s
    zn[z <= 0] = 0
    return zn
