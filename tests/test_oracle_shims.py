"""Independent second opinion on the PARITY-UNPINNED shims (oracle/shims: torchsparse v2.0.0 / spconv restatements):
every sparse primitive is compared with a dense ATen op (F.conv3d, F.conv_transpose3d, 3-D F.grid_sample) on the
scattered dense volume, and the back-projection arithmetic with a hand-computable case.  CPU only."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import restate  # noqa: F401  (puts oracle/shims on sys.path)
from torchsparse import SparseTensor
from torchsparse.nn import functional as TF
from torchsparse.nn.utils import get_kernel_offsets
import spconv.pytorch as spconv


def _random_sites(D=9, frac=0.35, seed=0):
    g = torch.Generator().manual_seed(seed)
    occ = torch.rand(D, D, D, generator=g) < frac
    xyz = torch.nonzero(occ)
    perm = torch.randperm(len(xyz), generator=g)
    return occ, xyz[perm]


def _dense(xyz, feats, D):
    vol = torch.zeros(1, feats.shape[1], D, D, D)
    vol[0, :, xyz[:, 0], xyz[:, 1], xyz[:, 2]] = feats.t()
    return vol


def test_hash_matches_fnv1a_reference_values():
    # independent scalar implementation of the published kernel (hash_cuda.cu)
    def h(c):
        v = 14695981039346656037
        for x in c:
            v ^= x & 0xFFFFFFFF
            v = (v * 1099511628211) & 0xFFFFFFFFFFFFFFFF
        return (v >> 60) ^ (v & 0x0FFFFFFFFFFFFFFF)
    coords = torch.tensor([[0, 0, 0, 0], [1, 2, 3, 0], [-1, -7, 5, 1], [100, -100, 31, 0]], dtype=torch.int32)
    got = TF.sphash(coords).tolist()
    assert got == [h(c) for c in coords.tolist()]
    off = get_kernel_offsets(3, 1)
    got_k = TF.sphash(coords, off)
    assert got_k.shape == (27, 4)
    assert got_k[0, 1].item() == h([1 - 1, 2 - 1, 3 - 1, 0])      # first offset is (-1,-1,-1): x-inner / z-outer


def test_kernel_offset_orderings():
    assert get_kernel_offsets(3, 1)[:4].tolist() == [[-1, -1, -1], [0, -1, -1], [1, -1, -1], [-1, 0, -1]]   # z-outer, x-inner
    assert get_kernel_offsets(2, 2).tolist() == [[0, 0, 0], [0, 0, 2], [0, 2, 0], [0, 2, 2], [2, 0, 0], [2, 0, 2], [2, 2, 0], [2, 2, 2]]


@pytest.mark.parametrize("cin,cout", [(5, 7), (16, 4)])
def test_conv3d_k3_equals_dense_conv(cin, cout):
    D = 9
    occ, xyz = _random_sites(D)
    g = torch.Generator().manual_seed(1)
    feats = torch.randn(len(xyz), cin, generator=g)
    kernel = torch.randn(27, cin, cout, generator=g)
    st = SparseTensor(feats, torch.cat([xyz, torch.zeros(len(xyz), 1, dtype=xyz.dtype)], 1).int(), 1)
    out = TF.conv3d(st, kernel, None, 3, 1, 1)
    w = torch.zeros(cout, cin, 3, 3, 3)
    for k, (dx, dy, dz) in enumerate(get_kernel_offsets(3, 1).tolist()):
        w[:, :, dx + 1, dy + 1, dz + 1] = kernel[k].t()
    dense = F.conv3d(_dense(xyz, feats, D), w, padding=1)
    want = dense[0, :, xyz[:, 0], xyz[:, 1], xyz[:, 2]].t()
    assert torch.allclose(out.F, want, atol=1e-4)


def test_strided_and_transposed_conv_equal_dense():
    D, cin, cmid = 8, 6, 5
    occ, xyz = _random_sites(D, frac=0.4, seed=3)
    g = torch.Generator().manual_seed(2)
    feats = torch.randn(len(xyz), cin, generator=g)
    k_down = torch.randn(8, cin, cmid, generator=g)
    k_up = torch.randn(8, cmid, cin, generator=g)
    st = SparseTensor(feats, torch.cat([xyz, torch.zeros(len(xyz), 1, dtype=xyz.dtype)], 1).int(), 1)
    st.cmaps.setdefault(st.stride, st.coords)
    down = TF.conv3d(st, k_down, None, 2, 2, 1)
    assert down.s == (2, 2, 2)
    # output sites: unique(floor(c/2)*2), sorted by (b,x,y,z)
    want_sites = torch.unique(torch.div(xyz, 2, rounding_mode="floor") * 2, dim=0)
    assert torch.equal(down.C[:, :3].long(), want_sites)
    w = torch.zeros(cmid, cin, 2, 2, 2)
    for k, (dx, dy, dz) in enumerate(get_kernel_offsets(2, 1).tolist()):
        w[:, :, dx, dy, dz] = k_down[k].t()
    dense = F.conv3d(_dense(xyz, feats, D), w, stride=2)
    cs = (down.C[:, :3] // 2).long()
    assert torch.allclose(down.F, dense[0, :, cs[:, 0], cs[:, 1], cs[:, 2]].t(), atol=1e-4)
    up = TF.conv3d(down, k_up, None, 2, 2, 1, transposed=True)
    assert up.s == (1, 1, 1) and torch.equal(up.C, st.C)
    wt = torch.zeros(cmid, cin, 2, 2, 2)
    for k, (dx, dy, dz) in enumerate(get_kernel_offsets(2, 1).tolist()):
        wt[:, :, dx, dy, dz] = k_up[k]
    dvol = torch.zeros(1, cmid, D // 2, D // 2, D // 2)
    dvol[0, :, cs[:, 0], cs[:, 1], cs[:, 2]] = down.F.t()
    dense_up = F.conv_transpose3d(dvol, wt, stride=2)
    assert torch.allclose(up.F, dense_up[0, :, xyz[:, 0], xyz[:, 1], xyz[:, 2]].t(), atol=1e-4)


def test_spdownsample_truncates_toward_zero_for_negative_coords():
    c = torch.tensor([[-3, -1, 0, 0], [-2, 1, 3, 0], [5, -5, 2, 0]], dtype=torch.int32)
    out = TF.spdownsample(c, 2, 2, 1)
    assert sorted(out.tolist()) == sorted([[-2, 0, 0, 0], [-2, 0, 2, 0], [4, -4, 2, 0]])


def test_devoxelize_equals_trilinear_grid_sample_on_full_grid():
    D, c = 6, 4
    g = torch.Generator().manual_seed(5)
    xyz = torch.stack(torch.meshgrid(*[torch.arange(D)] * 3, indexing="ij"), -1).view(-1, 3)
    feats = torch.randn(len(xyz), c, generator=g)
    st = SparseTensor(feats, torch.cat([xyz, torch.zeros(len(xyz), 1, dtype=xyz.dtype)], 1).int(), 1)
    pts = torch.rand(200, 3, generator=g) * (D - 1.001)
    pts4 = torch.cat([pts, torch.zeros(200, 1)], 1)
    idx, w = restate.trilinear_taps(pts4, st)
    got = TF.spdevoxelize(st.F, idx, w)
    vol = _dense(xyz, feats, D)                                   # [1,C,Dx,Dy,Dz]
    grid = (pts / (D - 1) * 2 - 1)[:, [2, 1, 0]].view(1, 1, 1, -1, 3)   # grid_sample wants (z,y,x) -> (W,H,D) order
    want = F.grid_sample(vol, grid, mode="bilinear", align_corners=True).view(c, -1).t()
    assert torch.allclose(got, want, atol=1e-5)


def test_voxelize_is_mean_pooling():
    pts = torch.tensor([[0.1, 0.2, 0.3, 0], [0.9, 0.8, 0.7, 0], [1.5, 0.5, 0.5, 0], [-0.5, 0.5, 0.5, 0]])
    feat = torch.tensor([[1.0], [3.0], [10.0], [20.0]])
    st, scaled, iq, cnt = restate.voxelize_points(feat, pts, 1.0)
    assert st.C.shape[0] == 3 and sorted(st.F.view(-1).tolist()) == [2.0, 10.0, 20.0]
    assert sorted(map(tuple, st.C[:, :3].tolist())) == [(-1, 0, 0), (0, 0, 0), (1, 0, 0)]


def test_subm_conv_equals_dense_conv_at_active_sites():
    D, cin, cout = 7, 5, 3
    occ, xyz = _random_sites(D, seed=7)
    g = torch.Generator().manual_seed(7)
    feats = torch.randn(len(xyz), cin, generator=g)
    conv = spconv.SubMConv3d(cin, cout, 3)
    idx = torch.cat([torch.zeros(len(xyz), 1, dtype=xyz.dtype), xyz], 1).int()
    with torch.no_grad():
        out = conv(spconv.SparseConvTensor(feats, idx, (D, D, D), 1)).features
        w = conv.weight.permute(0, 4, 1, 2, 3)                    # [Cout,Cin,k,k,k]
        dense = F.conv3d(_dense(xyz, feats, D), w, conv.bias, padding=1)
    assert torch.allclose(out, dense[0, :, xyz[:, 0], xyz[:, 1], xyz[:, 2]].t(), atol=1e-4)


def test_back_projection_hand_case():
    """2 views, 2x2 feature map, 3 voxels: visibility / mean / variance by hand."""
    feats = torch.zeros(2, 1, 1, 2, 2)
    feats[0, 0, 0] = torch.tensor([[1.0, 2.0], [3.0, 4.0]])
    feats[1, 0, 0] = torch.tensor([[10.0, 20.0], [30.0, 40.0]])
    P = torch.eye(4).repeat(2, 1, 1, 1)                            # pixel = (x/z, y/z), depth = z
    P[1, 0, 0, 3] = 0.5                                            # view 1 shifted by half a pixel in x (at z = 1)
    coords = torch.tensor([[0, 0, 0, 1], [0, 1, 1, 1], [0, 0, 0, -1]], dtype=torch.int32)   # last one behind the cameras
    origin = torch.zeros(1, 3)
    r = restate.backproject(coords, origin, 1.0, feats, P, 0)
    assert r["count"].tolist() == [2.0, 1.0, 0.0]                  # voxel 1 leaves view 1 (x = 1.5 > W-1)
    assert r["mask"].t().tolist() == [[True, True], [True, False], [False, False]]
    assert torch.allclose(r["feat"].view(-1), torch.tensor([(1.0 + 15.0) / 2, 4.0, 0.0]))
    v = restate.backproject(coords, origin, 1.0, feats, P, 2, mode="meanvar")
    assert v["coords"].tolist() == [[0, 0, 0, 1]] and torch.allclose(v["feat"].view(-1), torch.tensor([49.0]))


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference tree (build container only)")
def test_two_reference_back_projections_agree_with_the_restatement():
    """Back_Project vs the legacy back_project (two independent reference implementations of the same maths) vs restate."""
    from oracle import ref_import
    from eprecon_b200 import synth
    ns = ref_import.load()
    inputs, fa, fb = synth.make_fragment(seed=1, image_hw=(240, 320), n_vox=(64, 64, 64))
    g = torch.stack(torch.meshgrid(*[torch.arange(0, 64, 4)] * 3, indexing="ij")).view(3, -1)
    coords = torch.cat([torch.zeros(1, g.shape[1], dtype=torch.long), g]).t().contiguous().int()
    feats = torch.stack([f[2] for f in fb])
    kr = inputs["proj_matrices"][:, :, 2].permute(1, 0, 2, 3).contiguous()
    origin = inputs["vol_origin_partial"]
    a = ns.occ.Back_Project(80)(coords, origin, 0.04, feats, kr, 2)
    b = ns.back_project(coords, origin, 0.04, feats, kr, 2)
    c = restate.backproject(coords, origin, 0.04, feats, kr, 2)
    assert torch.equal(a[4], b[2]) and torch.equal(a[4], c["count"])
    assert torch.equal(a[1], c["coords"]) and torch.equal(a[3], c["mask"])
    assert torch.allclose(a[0], b[0][:, :80], atol=1e-6) and torch.allclose(a[0], c["feat"], atol=1e-4)
    assert torch.allclose(b[0][:, 80:], restate.legacy_depth_channel(c["zbar"]), atol=1e-6)
