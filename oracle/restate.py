"""TEST INFRASTRUCTURE ONLY — self-contained CPU restatement of the EPRecon feature-volume hot path.

Functional style (explicit state dicts, no nn.Module mirror) so that it can travel to the GPU box,
where /root/reference does not exist.  Each function cites the reference lines it follows; the
restatement is pinned against the UNMODIFIED reference (run in the build container) by the fixtures in
tests/golden/, each committed with the script that made it:
  neucon_small.npz          NeuConNet.forward over oracle/shims (make_golden.py)
  mask3dformer_small.npz    the panoptic decoder + panoptic_inference, plain ATen: a direct pin (make_golden_mask3dformer.py)
  panoptic_fusion_small.npz GRUFusion(direct_substitute) + panoptic_fusion, plain ATen: a direct pin (make_golden_panoptic_fusion.py)
The third-party sparse-conv semantics underneath the first fixture (torchsparse v2.0.0 / spconv) are restated
from their published algorithms and are PARITY UNPINNED — see oracle/shims/torchsparse/__init__.py.

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


# =========================================================================== sparse primitives (shims)
from torchsparse import SparseTensor  # noqa: E402  (oracle/shims)
from torchsparse.nn import functional as TF  # noqa: E402
from torchsparse.nn.utils import get_kernel_offsets  # noqa: E402
import spconv.pytorch as spconv_shim  # noqa: E402


def _bn(x, sd, p, eps=1e-5):
    """BatchNorm1d with the statistics of the current batch (the reference evaluates in train mode, main.py:357)."""
    return F.batch_norm(x, None, None, sd[p + ".weight"], sd[p + ".bias"], True, 0.1, eps)


def _ln(x, sd, p, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], eps)


def _lin(x, sd, p):
    return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])


def voxelize_points(feat, pts, vres):
    """initial_voxelize (ops/torchsparse_utils.py:15-35): returns (SparseTensor, scaled points, idx_query, counts)."""
    scaled = torch.cat([(pts[:, :3] * 1) / vres, pts[:, -1].view(-1, 1)], 1)
    fl = torch.floor(scaled)
    h = TF.sphash(fl.int())
    uniq = torch.unique(h)
    iq = TF.sphashquery(h, uniq)
    cnt = TF.spcount(iq.int(), len(uniq))
    vc = torch.round(TF.spvoxelize(fl, iq, cnt)).int()
    st = SparseTensor(TF.spvoxelize(feat, iq, cnt), vc, 1)
    st.cmaps.setdefault(st.stride, st.coords)
    return st, scaled, iq, cnt


def trilinear_taps(scaled, st):
    """voxel_to_point lookup part (ops/torchsparse_utils.py:71-82): (idx [N,8], w [N,8]) for SparseTensor st."""
    s = st.s[0]
    off = get_kernel_offsets(2, st.s, 1)
    base = torch.cat([torch.floor(scaled[:, :3] / s).int() * s, scaled[:, -1].int().view(-1, 1)], 1)
    idx = TF.sphashquery(TF.sphash(base, off), TF.sphash(st.C))
    w = TF.calc_ti_weights(scaled, idx, scale=s).transpose(0, 1).contiguous()
    return idx.transpose(0, 1).contiguous(), w


def points_to_voxels(feat, scaled, st):
    """point_to_voxel (ops/torchsparse_utils.py:40-63)."""
    s = st.s[0]
    h = TF.sphash(torch.cat([torch.floor(scaled[:, :3] / s).int() * s, scaled[:, -1].int().view(-1, 1)], 1))
    iq = TF.sphashquery(h, TF.sphash(st.C))
    cnt = TF.spcount(iq.int(), st.C.shape[0])
    return st._like(TF.spvoxelize(feat, iq, cnt))


def _sconv(st, sd, p, ks=3, stride=1, transposed=False):
    return TF.conv3d(st, sd[p + ".kernel"], None, ks, stride, 1, transposed)


def _conv_bn_relu(st, sd, p, **kw):
    y = _sconv(st, sd, p + ".0", **kw)
    return y._like(F.relu(_bn(y.F, sd, p + ".1")))


def _res_block(st, sd, p):
    """ResidualBlock (models/modules.py:46-73)."""
    t = _conv_bn_relu(st, sd, p + ".net")
    u = _bn(_sconv(t, sd, p + ".net.3").F, sd, p + ".net.4")
    if (p + ".downsample.0.kernel") in sd:
        d = _bn(st.F.matmul(sd[p + ".downsample.0.kernel"]), sd, p + ".downsample.1")
    else:
        d = st.F
    return st._like(F.relu(u + d))


def _cat(a, b):
    out = a._like(torch.cat([a.F, b.F], 1))
    return out


def spvcnn(sd, p, feat, pts, vres):
    """SPVCNN.forward (models/modules.py:148-175).  feat [N,Cin], pts float [N,4]=(x,y,z,b) -> [N, cs4]."""
    x0, scaled, _, _ = voxelize_points(feat, pts, vres)
    x0 = _conv_bn_relu(x0, sd, p + ".stem")
    idx1, w1 = trilinear_taps(scaled, x0)
    z0 = TF.spdevoxelize(x0.F, idx1, w1)
    x1 = points_to_voxels(z0, scaled, x0)
    x1 = _conv_bn_relu(x1, sd, p + ".stage1.0.net", ks=2, stride=2)
    x1 = _res_block(_res_block(x1, sd, p + ".stage1.1"), sd, p + ".stage1.2")
    x2 = _conv_bn_relu(x1, sd, p + ".stage2.0.net", ks=2, stride=2)
    x2 = _res_block(_res_block(x2, sd, p + ".stage2.1"), sd, p + ".stage2.2")
    idx4, w4 = trilinear_taps(scaled, x2)
    pt0 = F.relu(_bn(_lin(z0, sd, p + ".point_transforms.0.0"), sd, p + ".point_transforms.0.1"))
    z1 = TF.spdevoxelize(x2.F, idx4, w4) + pt0
    y3 = points_to_voxels(z1, scaled, x2)
    y3 = _conv_bn_relu(y3, sd, p + ".up1.0.net", ks=2, stride=2, transposed=True)
    y3 = _res_block(_res_block(_cat(y3, x1), sd, p + ".up1.1.0"), sd, p + ".up1.1.1")
    y4 = _conv_bn_relu(y3, sd, p + ".up2.0.net", ks=2, stride=2, transposed=True)
    y4 = _res_block(_res_block(_cat(y4, x0), sd, p + ".up2.1.0"), sd, p + ".up2.1.1")
    pt1 = F.relu(_bn(_lin(z1, sd, p + ".point_transforms.1.0"), sd, p + ".point_transforms.1.1"))
    return TF.spdevoxelize(y4.F, idx1, w1) + pt1


def _sconv3d(sd, p, feat, pts, vres, taps=None):
    """SConv3d.forward (models/modules.py:189-197); `taps` = cached (idx, w) to reuse (the convr quirk)."""
    st, scaled, _, _ = voxelize_points(feat, pts, vres)
    y = _sconv(st, sd, p + ".net")
    if taps is None:
        taps = trilinear_taps(scaled, y)
    out = TF.spdevoxelize(y.F, taps[0], taps[1]) + _lin(feat, sd, p + ".point_transforms.0")
    return out, scaled, taps


def convgru(sd, p, h, x, pts, vres):
    """ConvGRU.forward (models/modules.py:207-222) including the shared-PointTensor side effects: convz rebinds
    hx.C to pts/vres and caches stride-1 taps; convr therefore voxelises pts/vres^2 and devoxelises with convz's
    cached taps (ops/torchsparse_utils.py:33,69-70,97-99); convq sees a fresh PointTensor (original pts)."""
    hx = torch.cat([h, x], 1)
    z_pre, scaled1, taps1 = _sconv3d(sd, p + ".convz", hx, pts, vres)
    r_pre, _, _ = _sconv3d(sd, p + ".convr", hx, scaled1, vres, taps=taps1)
    z, r = torch.sigmoid(z_pre), torch.sigmoid(r_pre)
    q_pre, _, _ = _sconv3d(sd, p + ".convq", torch.cat([r * h, x], 1), pts, vres)
    return (1 - z) * h + z * torch.tanh(q_pre)


def linear4x(sd, p, x):
    """Linear4xTrans.forward (models/modules.py:298-311)."""
    o = F.relu(_ln(_lin(x, sd, p + ".linear1"), sd, p + ".norm1"))
    o = F.relu(_ln(_lin(o, sd, p + ".linear2"), sd, p + ".norm2"))
    o2 = _lin(o, sd, p + ".linear3")
    return o2 + o if sd[p + ".linear3.weight"].shape[0] == sd[p + ".linear3.weight"].shape[1] else o2


# ================================================================================= occupancy initialisation
def _conv2d_block(sd, p, x):
    y = F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding="same")
    return F.relu(F.batch_norm(y, None, None, sd[p + ".bn.weight"], sd[p + ".bn.bias"], True, 0.1, 1e-5))


def _elan2d(sd, p, x):
    f1, f2 = _conv2d_block(sd, p + ".conv1", x), _conv2d_block(sd, p + ".conv2", x)
    c3 = _conv2d_block(sd, p + ".conv3", f2)
    c4 = _conv2d_block(sd, p + ".conv4", c3)
    c5 = _conv2d_block(sd, p + ".conv5", c4)
    c6 = _conv2d_block(sd, p + ".conv6", c5)
    return _conv2d_block(sd, p + ".conv7", torch.cat([f1, f2, c3, c4, c5, c6], 1))


def _fusion_block(sd, p, x):
    bn2 = lambda y, q: F.batch_norm(y, None, None, sd[q + ".weight"], sd[q + ".bias"], True, 0.1, 1e-5)  # noqa: E731
    o = F.relu(bn2(F.conv2d(x, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding="same"), p + ".bn1"))
    o = F.relu(bn2(F.conv2d(o, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding="same"), p + ".bn2"))
    return _elan2d(sd, p + ".ELAN", o)


def _conv2d_res(sd, p, x):
    y = F.relu(F.conv2d(x, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding="same")) + x
    return F.batch_norm(y, None, None, sd[p + ".bn.weight"], sd[p + ".bn.bias"], True, 0.1, 1e-5)


def feat_fusion_pre(sd, p, f1x, f2x, f4x):
    """Occupancy_Initialization.feat_fusion_pre (models/occupancy_initialization.py:41-58)."""
    a = F.interpolate(_fusion_block(sd, p + ".self_fusion_1x", f1x), scale_factor=2, mode="bilinear")
    b = _fusion_block(sd, p + ".self_fusion_2x", f2x)
    c = F.avg_pool2d(_fusion_block(sd, p + ".self_fusion_4x", f4x), 2)
    x = _conv2d_block(sd, p + ".fusion_down", torch.cat([a, b, c], 1))
    for k in (1, 2, 3, 4):
        x = _conv2d_res(sd, f"{p}.post_fusion_{k}", x)
    return x


def _subm(sd, p, x, coords, shape, ks):
    conv = spconv_shim.SubMConv3d(sd[p + ".weight"].shape[-1], sd[p + ".weight"].shape[0], ks)
    conv.weight.data, conv.bias.data = sd[p + ".weight"], sd[p + ".bias"]
    with torch.no_grad():
        return conv(spconv_shim.SparseConvTensor(x, coords, shape, 1)).features


def _subm_block(sd, p, x, coords, shape, ks):
    return F.relu(_ln(_subm(sd, p + ".conv", x, coords, shape, ks), sd, p + ".ln"))


def occupancy_initialization(sd, p, coords, origin, voxel_size, features_all, krcam, shape, stage, min_view_number):
    """Occupancy_Initialization.forward (models/occupancy_initialization.py:61-182), bs == 1."""
    f1x = torch.stack([f[2] for f in features_all])[:, 0]
    f2x = torch.stack([f[1] for f in features_all])[:, 0]
    f4x = torch.stack([f[0] for f in features_all])[:, 0]
    fused = feat_fusion_pre(sd, p, f1x, f2x, f4x).unsqueeze(1)
    bp = backproject(coords, origin, voxel_size, fused, krcam, min_view_number, mode="meanvar", min_valid=1000)
    if bp is None:
        return None
    sc = (bp["coords"] / (2 ** (2 - stage))).to(coords.dtype)
    sc[:, 0] = 0
    x = _bn(bp["feat"], sd, p + ".norm0")
    e = p + ".similary_1"
    f1, f2 = _subm_block(sd, e + ".conv1", x, sc, shape, 1), _subm_block(sd, e + ".conv2", x, sc, shape, 1)
    c3 = _subm_block(sd, e + ".conv3", f2, sc, shape, 3)
    c4 = _subm_block(sd, e + ".conv4", c3, sc, shape, 3)
    c5 = _subm_block(sd, e + ".conv5", c4, sc, shape, 3)
    c6 = _subm_block(sd, e + ".conv6", c5, sc, shape, 3)
    x = _subm_block(sd, e + ".conv7", torch.cat([f1, f2, c3, c4, c5, c6], -1), sc, shape, 1)
    for k in (1, 2, 3):
        y = F.relu(_subm(sd, f"{p}.subm{k}.sparsesubmconv3d", x, sc, shape, 3)) + x
        x = _ln(y, sd, f"{p}.norm{k}")
    y = _bn(_subm(sd, p + ".subm4.sparsesubmconv3d", x, sc, shape, 3), sd, p + ".norm4")
    return {"occ": y, "coords": bp["coords"], "count": bp["count"], "src": bp["src"], "var": bp["feat"]}


def init_prune(occ_logit, count, shape_init, min_view_number=2, thr=0.3):
    """models/neucon_network.py:264,298-318 for bs == 1: -> int64 [n0,4] (b, x*4, y*4, z*4)."""
    sel = occ_logit.sigmoid().squeeze(-1) > thr
    vol = torch.zeros(shape_init, dtype=torch.bool)
    vol[count.view(shape_init) >= min_view_number] = sel
    v = F.max_pool3d(vol.unsqueeze(0).float(), 2).squeeze(0)
    ones = torch.ones((1, 1, 3, 3, 3))
    v = (F.conv3d(v.float()[None, None], ones, padding=1) == 27).squeeze()
    v = (F.conv3d(v.float()[None, None], ones, padding=1) >= 1).squeeze()
    v = (F.conv3d(v.float()[None, None], ones, padding=1) >= 1).squeeze()
    idx = torch.nonzero(v)
    return torch.cat([torch.zeros(len(idx), 1, dtype=idx.dtype), idx * 4], 1)


# =========================================================================================== GRU fusion
def aligned_points(coords, origin, voxel_size, w2ac, zero_batch=False):
    """world -> aligned-camera points (models/neucon_network.py:387-398; models/gru_fusion.py:331-337), numpy fp32,
    one rounding per op, dot products left to right (the order the CUDA kernel fixes)."""
    c = coords.cpu().numpy()
    b = c[:, 0].astype(np.int64)
    org = origin.cpu().numpy().astype(f32)
    R = w2ac.cpu().numpy().astype(f32)[b]
    vs = f32(voxel_size)
    w = [c[:, 1 + k].astype(f32) * vs + org[b, k] for k in range(3)]
    out = np.zeros((c.shape[0], 4), dtype=f32)
    for j in range(3):
        out[:, j] = ((R[:, j, 0] * w[0] + R[:, j, 1] * w[1]) + R[:, j, 2] * w[2]) + R[:, j, 3]
    out[:, 3] = 0 if zero_batch else c[:, 0].astype(f32)
    return torch.from_numpy(out)


class FusionState:
    """Per-scene recurrent state of GRUFusion (models/gru_fusion.py:31-38,59-65)."""

    def __init__(self):
        self.scene = [None] * 3
        self.origin = [None] * 3
        self.C = [None] * 3
        self.F = [None] * 3
        self.tC = [None] * 3
        self.tF = [None] * 3
        self.I = torch.zeros(0, dtype=torch.int32)   # global instance / semantic id per row of C[n_scales] (direct substitute)
        self.S = torch.zeros(0, dtype=torch.int32)


def _dense(locs, vals, dims, c, default):
    d = torch.full([dims[0], dims[1], dims[2], c], float(default), dtype=vals.dtype)
    if locs.shape[0] > 0:
        d[locs[:, 0], locs[:, 1], locs[:, 2]] = vals
    return d


def _common_rows(a, b):
    """Number of coincident points of two coordinate lists with unique rows (compute_overlap's `distances == 0`)."""
    key = lambda x: ((x[:, 0].long() + 4096) * 8192 + (x[:, 1].long() + 4096)) * 8192 + (x[:, 2].long() + 4096)  # noqa: E731
    return int(np.intersect1d(key(a).numpy(), key(b).numpy()).size)


def panoptic_fusion(state, scale, valid, rel, seg_u, info, upd, stuff_max=2, overlap_threshold=0.05):
    """GRUFusion.panoptic_fusion + compute_overlap (models/gru_fusion.py:116-193): relabel the fragment's segments with
    scene-level instance ids -- a thing segment takes the id of the first (ascending id) scene instance of its class
    that overlaps it with IoU > 0.05, else a new id; stuff segments take their class id."""
    cur = upd + rel
    gI, gS = state.I[valid], state.S[valid]
    max_id = max(int(state.I.max()), stuff_max) if len(state.I) > 0 else stuff_max
    new_i, new_s = torch.zeros_like(seg_u), torch.zeros_like(seg_u)
    inc = 1
    for i, d in enumerate(info):
        cls, sel = int(d["category_id"]), seg_u == i + 1
        if not d["isthing"]:
            new_i[sel], new_s[sel] = cls, cls
            continue
        matched = False
        if bool((gS == cls).any()):
            for ins in torch.unique(gI[gS == cls]).tolist():
                a, b = state.C[scale][state.I == ins], cur[sel]
                inter = _common_rows(a, b)
                union = len(a) + len(b) - inter
                if union > 0 and f32(inter) / f32(union) > f32(overlap_threshold):
                    new_i[sel], new_s[sel] = int(ins), cls
                    matched = True
                    break
        if not matched:
            new_i[sel], new_s[sel] = max_id + inc, cls
            inc += 1
    return new_i, new_s


def gru_fusion(state, sd, cfg, coords, values_in, inputs, scale, ch_voxel, direct_substitute=False, panoptic_info=None):
    """GRUFusion.forward (models/gru_fusion.py:259-394) for bs == 1, FUSION.FULL.  With `panoptic_info` (direct
    substitute only: {'panoptic_seg': [seg int32 [N], segments_info]}) the scene instance / semantic ids are fused too.
    Returns (coords int64 [U,4], values [U,C], tsdf_target [U,1], occ_target [U,1])."""
    interval = 2 ** (cfg.N_LAYER - scale - 1)
    scene = inputs["scene"][0]
    if state.scene[scale] is None or scene != state.scene[scale]:
        state.scene[scale] = scene
        state.origin[scale] = inputs["vol_origin"][0]
        c = 1 if direct_substitute else values_in.shape[1]
        state.C[scale], state.F[scale] = torch.zeros(0, 3, dtype=torch.long), torch.zeros(0, c)
        state.tC[scale], state.tF[scale] = torch.zeros(0, 3, dtype=torch.long), torch.zeros(0, 1)
        state.I, state.S = torch.zeros(0, dtype=torch.int32), torch.zeros(0, dtype=torch.int32)
    origin = inputs["vol_origin_partial"][0]
    voxel_size = cfg.VOXEL_SIZE * interval
    rel = ((origin - state.origin[scale]) / voxel_size).long()
    coords_b = torch.div(coords[:, 1:].long(), interval, rounding_mode="floor")
    dims = [int(n) // 2 ** (cfg.N_LAYER - scale - 1) for n in cfg.N_VOX]
    dim = torch.tensor(dims)
    c = values_in.shape[1]
    init = 1 if direct_substitute else 0
    gC = state.C[scale] - rel
    valid = ((gC < dim) & (gC >= 0)).all(-1)
    gvol = _dense(gC[valid], state.F[scale][valid], dims, c, init)
    cvol = _dense(coords_b, values_in, dims, c, init)
    if direct_substitute:
        upd = torch.nonzero((gvol.abs() < 1).any(-1) | (cvol.abs() < 1).any(-1))
    else:
        upd = torch.nonzero((gvol != 0).any(-1) | (cvol != 0).any(-1))
    lvl = cfg.N_LAYER - scale - 1
    has_gt = "occ_list" in inputs
    if has_gt:
        occ_t = inputs["occ_list"][lvl][0]
        tsdf_t = inputs["tsdf_list"][lvl][0][occ_t]
        tC = state.tC[scale] - rel
        valid_t = ((tC < dim) & (tC >= 0)).all(-1)
        tvol = _dense(torch.cat([tC[valid_t], torch.nonzero(occ_t)])[:, :3],
                      torch.cat([state.tF[scale][valid_t], tsdf_t.unsqueeze(-1)]), dims, 1, 1)
    values = cvol[upd[:, 0], upd[:, 1], upd[:, 2]]
    gvalues = gvol[upd[:, 0], upd[:, 1], upd[:, 2]]
    tsdf_target = tvol[upd[:, 0], upd[:, 1], upd[:, 2]] if has_gt else None
    occ_target = tsdf_target.abs() < 1 if has_gt else None
    if direct_substitute and panoptic_info is not None:
        # gru_fusion.py:352-363: the fragment's segment ids at the union sites (0 where only the scene has a voxel)
        seg, info = panoptic_info["panoptic_seg"]
        svol = torch.zeros(dims, dtype=seg.dtype)
        svol[coords_b[:, 0], coords_b[:, 1], coords_b[:, 2]] = seg
        new_i, new_s = panoptic_fusion(state, scale, valid, rel, svol[upd[:, 0], upd[:, 1], upd[:, 2]], info, upd)
        state.I = torch.cat([state.I[valid == False], new_i.to(torch.int32)])  # noqa: E712
        state.S = torch.cat([state.S[valid == False], new_s.to(torch.int32)])  # noqa: E712
    if not direct_substitute:
        cv = ch_voxel[scale]
        vres = cfg.VOXEL_SIZE * 2 ** (len(cfg.THRESHOLDS) - 1 - scale)
        upd0 = torch.cat([torch.zeros(len(upd), 1, dtype=upd.dtype), upd], 1)
        pts = aligned_points(upd0, origin.view(1, 3), voxel_size, inputs["world_to_aligned_camera"][:1], zero_batch=True)
        vv = convgru(sd, f"gru_fusion.fusion_nets_voxel.{scale}", gvalues[:, :cv], values[:, :cv], pts, vres)
        vi = convgru(sd, f"gru_fusion.fusion_nets_img.{scale}", gvalues[:, cv:], values[:, cv:], pts, vres)
        values = torch.cat([vv, vi], -1)
    state.F[scale] = torch.cat([state.F[scale][valid == False], values])  # noqa: E712
    state.C[scale] = torch.cat([state.C[scale][valid == False], upd + rel])  # noqa: E712
    if has_gt:
        tv = tvol.squeeze(-1)
        state.tF[scale] = torch.cat([state.tF[scale][valid_t == False], tv[tv.abs() < 1].unsqueeze(-1)])  # noqa: E712
        state.tC[scale] = torch.cat([state.tC[scale][valid_t == False], torch.nonzero(tv.abs() < 1) + rel])  # noqa: E712
    out_c = torch.cat([torch.zeros(len(upd), 1, dtype=upd.dtype), upd * interval], 1)
    return out_c, values, tsdf_target, occ_target


# ============================================================================================ NeuConNet
def upsample8(pre_feat, pre_coords, interval):
    """NeuConNet.upsample (models/neucon_network.py:193-214)."""
    pos = [[1], [2], [3], [1, 2], [1, 3], [2, 3], [1, 2, 3]]
    up_c = pre_coords.unsqueeze(1).repeat(1, 8, 1)
    for k, cols in enumerate(pos):
        for col in cols:
            up_c[:, k + 1, col] += interval
    up_f = pre_feat.unsqueeze(1).expand(-1, 8, -1).reshape(-1, pre_feat.shape[1])
    return up_f, up_c.view(-1, 4)


def neucon_forward(sd, cfg, features, features_b, inputs, state, trace=None, teacher=None, max_level=None):
    """NeuConNet.forward, TSDF path, bs == 1 (models/neucon_network.py:230-511).  `trace` (dict) receives every stage's
    tensors; `teacher` (a previous trace) overrides the data-dependent decisions (init selection, occupancy masks)
    so two implementations can be compared stage by stage on identical sparsity."""
    t = trace if trace is not None else {}
    n_scales = len(cfg.THRESHOLDS) - 1
    origin = inputs["vol_origin_partial"]
    axes = [torch.arange(0, n, 2) for n in cfg.N_VOX]
    g = torch.stack(torch.meshgrid(*axes, indexing="ij")).view(3, -1)
    shape_init = tuple(len(a) for a in axes)
    coords = torch.cat([torch.zeros(1, g.shape[1], dtype=torch.long), g]).t().contiguous().int()
    kr = inputs["proj_matrices"][:, :, 1].permute(1, 0, 2, 3).contiguous()
    init = occupancy_initialization(sd, "initialization", coords, origin, cfg.VOXEL_SIZE, features, kr, shape_init, 1, 2)
    if init is None:
        return None
    t["init"] = init
    sel = init_prune(init["occ"], init["count"], shape_init)
    t["init_selected"] = sel
    pre_feat = pre_coords = None
    channels = [96, 48, 24]
    out = {}
    for i in range(cfg.N_LAYER):
        interval, scale = 2 ** (n_scales - i), n_scales - i
        if i == 0:
            up_coords, minv = sel.int(), 2
        else:
            up_feat, up_coords = upsample8(pre_feat, pre_coords, interval)
            minv = 0
        feats = torch.stack([f[scale] for f in features_b])
        kr = inputs["proj_matrices"][:, :, scale].permute(1, 0, 2, 3).contiguous()
        bp = backproject(up_coords, origin, cfg.VOXEL_SIZE, feats, kr, minv)
        if bp is None:
            return None
        volume, up_coords = bp["feat"], bp["coords"]
        feat = torch.cat([volume, up_feat[bp["count"] >= minv]], 1) if i != 0 else volume
        pts = aligned_points(up_coords, origin, cfg.VOXEL_SIZE, inputs["world_to_aligned_camera"])
        vres = cfg.VOXEL_SIZE * 2 ** (n_scales - i)
        fv = spvcnn(sd, f"sp_convs.{i}", feat, pts, vres)
        feat_all = torch.cat([fv, volume], -1)
        t[f"l{i}_pre_gru"] = {"coords": up_coords, "feat_in": feat, "pts": pts, "spvcnn": fv}
        up_coords, feat_all, tsdf_target, occ_target = gru_fusion(state, sd, cfg, up_coords, feat_all, inputs, i, channels)
        fv = feat_all[:, :channels[i]]
        tsdf = linear4x(sd, f"tsdf_preds.{i}", fv)
        occ = linear4x(sd, f"occ_preds.{i}", fv)
        occupancy = occ.squeeze(1) > cfg.THRESHOLDS[i]
        if teacher is not None:
            occupancy = teacher[f"l{i}"]["occupancy"]
        t[f"l{i}"] = {"coords": up_coords, "feat_all": feat_all, "tsdf": tsdf, "occ": occ, "occupancy": occupancy,
                      "occ_target": occ_target}
        num = int(occupancy.sum())
        if num < 500 or num > cfg.TRAIN_NUM_SAMPLE[i] * 1.5:
            return None
        assert num <= cfg.TRAIN_NUM_SAMPLE[i], "keep synthetic occupancy under the caps (np.random drop not restated)"
        if occ_target[occupancy].sum() == 0:
            return None
        pre_coords = up_coords[occupancy]
        pre_feat = torch.cat([fv[occupancy], tsdf[occupancy], occ[occupancy]], 1)
        if i == cfg.N_LAYER - 1 or (max_level is not None and i >= max_level):
            out["coords"], out["tsdf"] = pre_coords, tsdf[occupancy]
            out["level"] = i
            break
    return out


# ======================================================================= panoptic feature preparation (next row)
def _subm_residual(sd, p, x, coords, shape):
    """SparseConv3d_Residual.forward (models/modules.py:476-482)."""
    y = F.relu(_subm(sd, p + ".SConv3d.sparsesubmconv3d", x, coords, shape, 3))
    return _ln(x + y, sd, p + ".norm")


def panoptic_prepare(sd, cfg, trace, chunk=2048):
    """models/neucon_network.py:516-560, bs == 1: level alignment by coordinate-row equality (the reference's broadcast
    compare, chunked here to bound memory), per-level Linear4xTrans, three residual SubM convs for the mask features."""
    coords = [trace[f"l{i}"]["coords"][trace[f"l{i}"]["occupancy"]] for i in range(3)]
    feats = [trace[f"l{i}"]["feat_all"][trace[f"l{i}"]["occupancy"]] for i in range(3)]

    def member(a, b):  # rows of a that occur in b
        out = torch.zeros(len(a), dtype=torch.bool)
        for s0 in range(0, len(a), chunk):
            blk = a[s0:s0 + chunk]
            hit = torch.zeros(len(blk), dtype=torch.bool)
            for t0 in range(0, len(b), 8192):
                hit |= (blk.unsqueeze(1) == b[t0:t0 + 8192].unsqueeze(0)).all(2).any(1)
            out[s0:s0 + chunk] = hit
        return out

    down = torch.unique(torch.cat([coords[2][:, :1], torch.floor_divide(coords[2][:, 1:], 2) * 2], 1), dim=0)
    m1 = member(coords[1], down)
    down_b = torch.unique(torch.cat([down[:, :1], torch.floor_divide(down[:, 1:], 4) * 4], 1), dim=0)
    m0 = member(coords[0], down_b)
    coords[1], feats[1] = coords[1][m1], feats[1][m1]
    coords[0], feats[0] = coords[0][m0], feats[0][m0]
    pf = [linear4x(sd, f"panoptic_preds.{p}", feats[p]) for p in range(3)]
    shape = tuple(int(n) for n in cfg.N_VOX)
    cz = torch.cat([torch.zeros_like(coords[2][:, :1]), coords[2][:, 1:]], 1).int()
    x = pf[2]
    for k in range(3):
        x = _subm_residual(sd, f"panoptic_feat_fusion.mask_feat_extraction_{k}", x, cz, shape)
    return {"coords": coords, "feats": pf, "mask_features": x}


# ======================================================================= panoptic decoder (next row)
# The decoder is plain ATen in the reference (models/mask3dformer.py, models/voxel_position_encoding.py): this
# restatement is pinned DIRECTLY against the unmodified reference modules (tests/golden/mask3dformer_small.npz).
N_HEADS = 8


def fourier_positions(sd, p, xyz, shape):
    """PositionEmbeddingCoordsSine.get_fourier_embeddings with normalize=True (models/voxel_position_encoding.py:116-146;
    called from mask3dformer.py:318-335 with input_range = [0, spatial shape]): xyz int [N,3] -> [N, d_pos]."""
    x = xyz.float()
    src_diff = torch.tensor([float(s) for s in shape], dtype=torch.float32)
    x = ((x - 0.0) * 1.0) / src_diff + 0.0        # shift_scale_points (:11-40) with dst range [0, 1]
    x = x * f32(2 * np.pi)
    proj = x @ sd[p + ".gauss_B"]
    return torch.cat([proj.sin(), proj.cos()], 1)


def _mha(sd, p, query, key, value, blocked=None, nheads=N_HEADS):
    """nn.MultiheadAttention forward (batch of one, no dropout): query [Q,E], key/value [N,E]; blocked bool [Q,N]
    (True = may not attend, the same for every head, mask3dformer.py:444)."""
    E = query.shape[1]
    W, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    q = F.linear(query, W[:E], b[:E])
    k = F.linear(key, W[E:2 * E], b[E:2 * E])
    v = F.linear(value, W[2 * E:], b[2 * E:])
    hd = E // nheads
    q = q.view(-1, nheads, hd).transpose(0, 1)           # [h,Q,hd]
    k = k.view(-1, nheads, hd).transpose(0, 1)
    v = v.view(-1, nheads, hd).transpose(0, 1)
    s = (q @ k.transpose(1, 2)) / float(np.sqrt(hd))
    if blocked is not None:
        s = s.masked_fill(blocked.unsqueeze(0), float("-inf"))
    a = torch.softmax(s, dim=-1)
    o = (a @ v).transpose(0, 1).reshape(-1, E)
    return F.linear(o, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


def nearest_fine_index(coarse_xyz, fine_xyz, chunk=64):
    """mask3dformer.py:361-368: `argmin(cdist(fine, coarse[None]), dim=1)` reduces over the FINE axis, i.e. for every
    coarse voxel the index of its nearest level-2 voxel (first index on ties; distances between integer voxels are
    exact in fp32, so ties are exact).  Brute force with integer distances."""
    fine = fine_xyz.to(torch.int64)
    out = torch.empty(len(coarse_xyz), dtype=torch.int64)
    big = int(len(fine)) + 1
    for s0 in range(0, len(coarse_xyz), chunk):
        c = coarse_xyz[s0:s0 + chunk].to(torch.int64)
        d2 = ((fine.unsqueeze(0) - c.unsqueeze(1)) ** 2).sum(-1)          # [chunk, N2]
        key = d2 * big + torch.arange(len(fine)).unsqueeze(0)
        out[s0:s0 + chunk] = key.min(dim=1).values % big
    return out


def _pred_heads(sd, p, output, mask_features, index):
    """forward_prediction_heads (mask3dformer.py:431-447); index None = level 2 (all voxels, :370)."""
    d = _ln(output, sd, p + ".decoder_norm")
    logits = _lin(d, sd, p + ".class_embed")
    e = F.relu(_lin(d, sd, p + ".mask_embed.layers.0"))
    e = F.relu(_lin(e, sd, p + ".mask_embed.layers.1"))
    e = _lin(e, sd, p + ".mask_embed.layers.2")
    masks = e @ mask_features.t()                                          # [Q, N2]
    sel = masks if index is None else masks[:, index]
    blocked = torch.sigmoid(sel) < 0.5
    return logits, masks, blocked


def mask3dformer(sd, p, feats, coords_xyz, mask_features, shape, n_layers=6):
    """MultiScaleMaskedTransformerDecoder.forward (mask3dformer.py:338-429), one fragment.
    feats[l] [N_l,48] (level-aligned panoptic features), coords_xyz[l] int [N_l,3], mask_features [N_2,48]."""
    pos = [fourier_positions(sd, p + ".pos_enc", coords_xyz[l], shape) for l in range(3)]
    src = [feats[l] + sd[p + ".level_embed.weight"][l].unsqueeze(0) for l in range(3)]
    index = [nearest_fine_index(coords_xyz[0], coords_xyz[2]), nearest_fine_index(coords_xyz[1], coords_xyz[2]), None]
    qpos = sd[p + ".query_embed.weight"]
    out = sd[p + ".query_feat.weight"]
    logits, masks, blocked = _pred_heads(sd, p, out, mask_features, index[0])
    aux = [(logits, masks)]
    for j in range(n_layers):
        l = j % 3
        blocked = blocked & ~blocked.all(dim=1, keepdim=True)             # a fully blocked query attends everywhere (:389)
        t2 = _mha(sd, f"{p}.transformer_cross_attention_layers.{j}.multihead_attn", out + qpos, src[l] + pos[l], src[l], blocked)
        out = _ln(out + t2, sd, f"{p}.transformer_cross_attention_layers.{j}.norm")
        t2 = _mha(sd, f"{p}.transformer_self_attention_layers.{j}.self_attn", out + qpos, out + qpos, out)
        out = _ln(out + t2, sd, f"{p}.transformer_self_attention_layers.{j}.norm")
        ff = f"{p}.transformer_ffn_layers.{j}"
        t2 = _lin(F.relu(_lin(out, sd, ff + ".linear1")), sd, ff + ".linear2")
        out = _ln(out + t2, sd, ff + ".norm")
        logits, masks, blocked = _pred_heads(sd, p, out, mask_features, index[(j + 1) % 3])
        aux.append((logits, masks))
    return {"pred_logits": logits, "pred_masks": masks, "aux": aux[:-1], "index": index[:2]}


def panoptic_inference(mask_cls, mask_pred, object_mask_threshold=0.3, thing_id=tuple(range(3, 21)), overlap_threshold=0.5):
    """mask3dformer.py:516-581: mask_cls [Q, classes+1], mask_pred [Q, N] logits -> (panoptic_seg int32 [N], segments)."""
    scores, labels = torch.softmax(mask_cls, -1).max(-1)
    prob = torch.sigmoid(mask_pred)
    keep = (labels != 0) & (scores > object_mask_threshold)
    seg = torch.zeros(mask_pred.shape[-1], dtype=torch.int32)
    info = []
    if int(keep.sum()) == 0:
        return seg, info
    ks, kc, km = scores[keep], labels[keep], prob[keep]
    winner = (ks.view(-1, 1) * km).argmax(0)
    next_id, stuff = 0, {}
    for k in range(len(kc)):
        cls = int(kc[k])
        thing = cls in thing_id
        area = int((winner == k).sum())
        orig = int((km[k] >= 0.5).sum())
        m = (winner == k) & (km[k] >= 0.5)
        if area > 0 and orig > 0 and int(m.sum()) > 0:
            if area / orig < overlap_threshold:
                continue
            if not thing:
                if cls in stuff:
                    seg[m] = stuff[cls]
                    continue
                stuff[cls] = next_id + 1
            next_id += 1
            seg[m] = next_id
            info.append({"id": next_id, "isthing": bool(thing), "category_id": cls})
    return seg, info
