"""CUDA sparse primitives and modules vs the CPU oracle (oracle/restate.py + oracle/shims)."""
import numpy as np
import pytest
import torch

from eprecon_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-3


def rel(a, b):
    a, b = a.detach().cpu().float(), b.detach().cpu().float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def shell_points(n_side=20, seed=0, spacing=0.16, bs=1):
    """Rotated regular grid shell (like voxel centres in the aligned-camera frame), some negative coordinates."""
    g = torch.Generator().manual_seed(seed)
    ax = torch.arange(n_side, dtype=torch.float32) * spacing
    p = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1).view(-1, 3)
    keep = ((p - p.mean(0)).norm(dim=1) > 0.8) & ((p - p.mean(0)).norm(dim=1) < 1.5)
    p = p[keep] - 1.0
    a = 0.3
    R = torch.tensor([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a) * np.cos(0.2), -np.sin(0.2)],
                      [0, np.sin(0.2), np.cos(0.2)]], dtype=torch.float32)
    p = p @ R.t() + 0.013
    b = torch.zeros(len(p), 1) if bs == 1 else torch.randint(0, bs, (len(p), 1), generator=g).float().sort(0)[0]
    return torch.cat([p, b], 1).contiguous()


def _row_key(c):
    c = c.long()
    return ((c[:, 3] * 4096 + c[:, 0] + 2048) * 4096 + c[:, 1] + 2048) * 4096 + c[:, 2] + 2048


def _same_rows(a, b):
    return torch.equal(torch.sort(_row_key(a))[0], torch.sort(_row_key(b))[0])


def _align(rows_mine, coords_mine, coords_want):
    """Reorder my rows (internal Z-order) into the oracle's row order by matching coordinates."""
    km, kw = _row_key(coords_mine), _row_key(coords_want)
    om, ow = torch.argsort(km), torch.argsort(kw)
    assert torch.equal(km[om], kw[ow])
    out = torch.empty_like(rows_mine)
    out[ow] = rows_mine[om]
    return out


def test_voxelize_taps_and_maps(cuda_lib):
    from oracle import restate
    from eprecon_b200 import sparse
    pts = shell_points()
    feat = torch.randn(len(pts), 12, generator=torch.Generator().manual_seed(1))
    st, scaled, iq, cnt = restate.voxelize_points(feat, pts, 0.16)
    pc = sparse.PointCloud(pts.cuda(), 0.16, order="hash")
    assert torch.equal(pc.vox.coords.cpu(), st.C)                       # same voxels in the same (hash-sorted) order
    assert torch.equal(pc.idx_query.cpu().long(), iq)
    assert torch.equal(pc.scaled.cpu(), scaled)
    vf = pc.voxelize(feat.cuda(), 12)
    assert rel(vf[:, :12], st.F) < 1e-6
    idx, w = restate.trilinear_taps(scaled, st)
    gi, gw = pc.taps(pc.vox)
    assert torch.equal(gi.cpu().long(), idx) and (gw.cpu() - w).abs().max() < 1e-6
    # strided maps: coarse voxel sets and the k2s2 / transposed tables agree with the shim's spdownsample + hash query
    from torchsparse.nn import functional as TF
    cc = TF.spdownsample(st.C, 2, 2, 1)
    v1, down, up = pc.vox.downsample()
    assert _same_rows(v1.coords.cpu(), cc)          # same coarse sites (row order is internal: Z-order here)
    cc2 = TF.spdownsample(cc, 2, 2, 2)
    v2, _, _ = v1.downsample()
    assert _same_rows(v2.coords.cpu(), cc2)


@pytest.mark.parametrize("cin,cout", [(12, 32), (138, 16), (32, 1), (160, 96)])
def test_sparse_conv_k3_and_strided(cuda_lib, cin, cout):
    from oracle import restate
    from eprecon_b200 import ops, sparse
    from torchsparse import SparseTensor
    from torchsparse.nn import functional as TF
    pts = shell_points(seed=2)
    g = torch.Generator().manual_seed(cin * 131 + cout)
    feat = torch.randn(len(pts), cin, generator=g)
    st, scaled, _, _ = restate.voxelize_points(feat, pts, 0.16)
    pc = sparse.PointCloud(pts.cuda(), 0.16, order="hash")
    fpad = torch.zeros(len(pts), ops.ceil4(cin))
    fpad[:, :cin] = feat
    x = pc.voxelize(fpad.cuda(), cin)
    W3 = torch.randn(27, cin, cout, generator=g) / (27 * cin) ** 0.5
    Wp = torch.zeros(27, cin, ops.ceil4(cout))
    Wp[:, :, :cout] = W3
    y, part = ops.spconv(x, cin, pc.vox.kmap_k3(), Wp.cuda(), cout, want_stats=True)
    want = TF.conv3d(st, W3, None, 3, 1, 1)
    assert rel(y[:, :cout], want.F) < 1e-5
    # fused BN statistics == column sums of the output
    s = part[:, 0].sum(0).cpu()
    assert torch.allclose(s, want.F.sum(0), rtol=1e-4, atol=1e-3)
    # k2s2 down and its transposed twin
    W2 = torch.randn(8, cin, cout, generator=g) / (8 * cin) ** 0.5
    W2p = torch.zeros(8, cin, ops.ceil4(cout))
    W2p[:, :, :cout] = W2
    v1, down, up = pc.vox.downsample()
    yd, _ = ops.spconv(x, cin, down, W2p.cuda(), cout)
    wd = TF.conv3d(st, W2, None, 2, 2, 1)
    assert rel(_align(yd[:, :cout].cpu(), v1.coords.cpu(), wd.C), wd.F) < 1e-5
    Wt = torch.randn(8, cout, cin, generator=g) / (8 * cout) ** 0.5
    Wtp = torch.zeros(8, cout, ops.ceil4(cin))
    Wtp[:, :, :cin] = Wt
    yd_pad = yd if yd.shape[1] == ops.ceil4(cout) else yd
    yu, _ = ops.spconv(yd_pad, cout, up, Wtp.cuda(), cin)
    wu = TF.conv3d(wd, Wt, None, 2, 2, 1, transposed=True)
    assert rel(yu[:, :cin], wu.F) < 1e-5


def test_spvcnn_levels(cuda_lib):
    from oracle import restate
    from eprecon_b200.modules import SPVCNN
    from eprecon_b200.tensor import PointTensor
    for lvl, (cin, cr, vres, n_side) in enumerate([(80, 1.0, 0.16, 18), (138, 0.5, 0.08, 24), (74, 0.25, 0.04, 30)]):
        net = SPVCNN(num_classes=1, in_channels=cin, pres=1, cr=cr, vres=vres, dropout=False)
        sd = {f"sp.{k}": v for k, v in synth.synthetic_state_dict(net, 3 + lvl).items()}
        pts = shell_points(n_side=n_side, seed=lvl, spacing=vres)
        feat = torch.randn(len(pts), cin, generator=torch.Generator().manual_seed(lvl))
        with torch.no_grad():
            want = restate.spvcnn(sd, "sp", feat, pts, vres)
        got = net.cuda()(PointTensor(feat.cuda(), pts.cuda()))
        assert got.shape == want.shape
        assert rel(got, want) < RTOL, (lvl, rel(got, want))


def test_convgru_with_reference_quirk(cuda_lib):
    from oracle import restate
    from eprecon_b200.modules import ConvGRU
    from eprecon_b200.tensor import PointTensor
    c, vres = 48, 0.08
    gru = ConvGRU(hidden_dim=c, input_dim=c, pres=1, vres=vres)
    sd = {f"g.{k}": v for k, v in synth.synthetic_state_dict(gru, 5).items()}
    pts = shell_points(n_side=22, seed=4, spacing=vres)
    pts[:, 3] = 0
    g = torch.Generator().manual_seed(9)
    h, x = torch.randn(len(pts), c, generator=g), torch.randn(len(pts), c, generator=g)
    h[::3] = 0  # global state absent on a third of the sites
    with torch.no_grad():
        want = restate.convgru(sd, "g", h, x, pts, vres)
    got = gru.cuda()(PointTensor(h.cuda(), pts.cuda()), PointTensor(x.cuda(), pts.cuda()))
    assert rel(got, want) < RTOL, rel(got, want)


def test_linear4x_heads(cuda_lib):
    from oracle import restate
    from eprecon_b200.modules import Linear4xTrans
    for cin, cout in [(96, 1), (24, 1), (48, 48)]:
        m = Linear4xTrans(cin, cout)
        sd = {f"h.{k}": v for k, v in synth.synthetic_state_dict(m, 2).items()}
        x = torch.randn(3001, cin, generator=torch.Generator().manual_seed(cin))
        want = restate.linear4x(sd, "h", x)
        got = m.cuda()(x.cuda())
        assert rel(got, want) < RTOL


def test_occupancy_initialization_and_prune(cuda_lib):
    from oracle import restate
    from eprecon_b200 import ops
    from eprecon_b200.neucon_network import NeuConNet
    n_vox = (64, 64, 64)
    cfg = synth.make_cfg(n_vox=n_vox)
    net = NeuConNet(cfg)
    sd = synth.synthetic_state_dict(net, 1)
    inputs, fa, fb = synth.make_fragment(seed=1, image_hw=(240, 320), n_vox=n_vox)
    axes = [torch.arange(0, n, 2) for n in n_vox]
    g = torch.stack(torch.meshgrid(*axes, indexing="ij")).view(3, -1)
    coords = torch.cat([torch.zeros(1, g.shape[1], dtype=torch.long), g]).t().contiguous().int()
    kr = inputs["proj_matrices"][:, :, 1].permute(1, 0, 2, 3).contiguous()
    with torch.no_grad():
        want = restate.occupancy_initialization(sd, "initialization", coords, inputs["vol_origin_partial"], 0.04, fa, kr,
                                                (32, 32, 32), 1, 2)
    net = net.cuda()
    fa_c = [[t.cuda() for t in f] for f in fa]
    got = net.initialization(coords.cuda(), inputs["vol_origin_partial"].cuda(), 0.04, fa_c, kr.cuda(), (32, 32, 32), 1, 2)
    assert torch.equal(got[1].cpu(), want["coords"]) and torch.equal(got[2].cpu(), want["count"])
    assert rel(net.initialization.last["feat"], want["var"]) < RTOL
    assert rel(got[0], want["occ"]) < RTOL
    # pruning on the ORACLE's logits (teacher forcing): bit-exact coarse selection
    sel = restate.init_prune(want["occ"], want["count"], (32, 32, 32))
    L = ops._L()
    fine = torch.zeros(32 ** 3, dtype=torch.uint8, device="cuda")
    occ_c = want["occ"].cuda().contiguous()
    ops._lib.check(L.ep_scatter_selected(occ_c.data_ptr(), 1, net.initialization.last["src"].data_ptr(), occ_c.shape[0],
                                         0.3, fine.data_ptr(), ops.stream_ptr()), "scatter")
    out = torch.empty((16 ** 3, 4), dtype=torch.int32, device="cuda")
    cnt = torch.empty(2, dtype=torch.int32, device="cuda")
    ops._lib.check(L.ep_init_prune(fine.data_ptr(), 1, 16, 4, out.data_ptr(), cnt.data_ptr(), ops.stream_ptr()), "prune")
    n0 = int(cnt[1].item())
    assert torch.equal(out[:n0].cpu().long(), sel)
