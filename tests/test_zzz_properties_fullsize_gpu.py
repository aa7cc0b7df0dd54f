"""Size-independent properties at BASELINE.json's FULL sizes (9 x 640 x 480, 96^3, ~210 k level-2 candidates), where the
CPU oracle is too slow to be the checker: linearity, order stability, implementation agreement, idempotence and
run-to-run determinism of the CUDA path (every kernel is meant to be deterministic: fixed-order reductions, no float
atomics)."""
import pytest
import torch

from eprecon_b200 import synth

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def _dev(obj):
    if torch.is_tensor(obj):
        return obj.cuda()
    if isinstance(obj, list):
        return [_dev(o) for o in obj]
    if isinstance(obj, dict):
        return {k: _dev(v) for k, v in obj.items()}
    return obj


@pytest.fixture(scope="module")
def level2(cuda_lib):
    """Level-2 candidates of the full-size synthetic fragment: the 8 children of every occupied 48^3 GT voxel."""
    inputs, fa, fb = synth.make_fragment(seed=1)
    par = torch.nonzero(inputs["occ_list"][1][0]).to(torch.int32) * 2
    offs = torch.tensor([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1]], dtype=torch.int32)
    xyz = (par.unsqueeze(1) + offs.unsqueeze(0)).reshape(-1, 3)
    coords = torch.cat([torch.zeros(len(xyz), 1, dtype=torch.int32), xyz], 1).contiguous().cuda()
    feats = torch.stack([f[0] for f in fb]).cuda()                              # [9,1,24,120,160]
    kr = inputs["proj_matrices"][:, :, 0].permute(1, 0, 2, 3).contiguous().cuda()
    return inputs, coords, feats, kr, inputs["vol_origin_partial"].cuda()


def test_back_projection_is_linear_in_the_features_and_ignores_them_for_masks(level2):
    from eprecon_b200 import ops
    _, coords, feats, kr, origin = level2
    assert coords.shape[0] > 200000
    g = torch.Generator(device="cuda").manual_seed(7)
    f2 = torch.randn(feats.shape, device="cuda", generator=g)
    ra = ops.backproject(coords, origin, 0.04, ops.to_nhwc(feats), kr, 0)
    rb = ops.backproject(coords, origin, 0.04, ops.to_nhwc(f2), kr, 0)
    rc = ops.backproject(coords, origin, 0.04, ops.to_nhwc(feats + 2.0 * f2), kr, 0)
    for r in (rb, rc):                                                          # geometry does not depend on the features
        assert torch.equal(r["coords"], ra["coords"]) and torch.equal(r["count"], ra["count"]) and torch.equal(r["vis"], ra["vis"])
    assert rel(rc["feat"], ra["feat"] + 2.0 * rb["feat"]) < 1e-5
    # masked mean: a constant map back-projects to that constant wherever a view sees the voxel, 0 elsewhere
    one = ops.backproject(coords, origin, 0.04, ops.to_nhwc(torch.ones_like(feats)), kr, 0)
    seen = one["count"] > 0
    assert (one["feat"][seen] - 1.0).abs().max().item() < 1e-5 and one["feat"][~seen].abs().max().item() == 0.0


def test_back_projection_compaction_is_stable_and_order_independent(level2):
    from eprecon_b200 import ops
    _, coords, feats, kr, origin = level2
    nhwc = ops.to_nhwc(feats)
    a = ops.backproject(coords, origin, 0.04, nhwc, kr, 2, want_src=True)
    src = a["src"].long()
    assert torch.equal(a["coords"], coords[src]) and bool((src[1:] > src[:-1]).all())   # survivors in input order
    assert torch.equal(a["count"][src] >= 2, torch.ones_like(src, dtype=torch.bool))
    perm = torch.randperm(coords.shape[0], device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    b = ops.backproject(coords[perm].contiguous(), origin, 0.04, nhwc, kr, 2, want_src=True)
    assert torch.equal(b["count"], a["count"][perm])
    # same voxel -> bit-identical row, wherever it sits in the input
    rows_b = torch.full((coords.shape[0],), -1, dtype=torch.long, device="cuda")
    rows_b[perm[b["src"].long()]] = torch.arange(b["src"].numel(), device="cuda")
    assert torch.equal(b["feat"][rows_b[src]], a["feat"])


def test_fused_and_three_pass_back_projection_agree_bitwise(level2):
    from eprecon_b200 import ops
    _, coords, feats, kr, origin = level2
    nhwc = ops.to_nhwc(feats)
    old = ops.BP_IMPL
    try:
        ops.BP_IMPL = "fused"
        a = ops.backproject(coords, origin, 0.04, nhwc, kr, 2)
        ops.BP_IMPL = "3pass"
        b = ops.backproject(coords, origin, 0.04, nhwc, kr, 2)
    finally:
        ops.BP_IMPL = old
    assert torch.equal(a["coords"], b["coords"]) and torch.equal(a["count"], b["count"]) and torch.equal(a["vis"], b["vis"])
    assert rel(a["feat"], b["feat"]) < 1e-6


def test_sparse_conv_full_size_linearity_and_implementation_agreement(level2):
    """27-offset submanifold map on the ~210 k level-2 sites: tcgen05 3xTF32 vs the fp32 FFMA kernel, linearity, and
    the fused BatchNorm statistics against column sums."""
    from eprecon_b200 import ops
    _, coords, _, _, _ = level2
    table = ops.HashTable(ops.coord_keys(coords, True))
    nbr = ops.kmap_build(coords, True, ops.kernel_offsets("subm3", 1, coords.device), table)
    m = coords.shape[0]
    assert bool((nbr[:, 13] == torch.arange(m, device="cuda", dtype=torch.int32)).all())    # centre tap = identity
    g = torch.Generator(device="cuda").manual_seed(11)
    cin, cout = 24, 24
    x1 = torch.randn(m, cin, device="cuda", generator=g)
    x2 = torch.randn(m, cin, device="cuda", generator=g)
    W = torch.randn(27, cin, cout, device="cuda", generator=g) / (27 * cin) ** 0.5
    old = ops.SPCONV_IMPL
    try:
        ops.SPCONV_IMPL = "ffma"
        y_ffma, _ = ops.spconv(x1, cin, nbr, W, cout)
        ops.SPCONV_IMPL = "tf32x3"
        y1, part = ops.spconv(x1, cin, nbr, W, cout, want_stats=True)
        y2, _ = ops.spconv(x2, cin, nbr, W, cout)
        y3, _ = ops.spconv(x1 - 0.5 * x2, cin, nbr, W, cout)
        y1b, _ = ops.spconv(x1, cin, nbr, W, cout)
    finally:
        ops.SPCONV_IMPL = old
    assert rel(y1, y_ffma) < 3e-5
    assert rel(y3, y1 - 0.5 * y2) < 3e-5
    assert torch.equal(y1, y1b)                                                   # run-to-run determinism
    assert torch.allclose(part[:, 0].sum(0), y1.sum(0), rtol=1e-3, atol=0.1)
    assert torch.allclose(part[:, 1].sum(0), (y1 ** 2).sum(0), rtol=1e-3, atol=0.1)


def test_scene_fusion_is_idempotent(level2):
    """Direct-substitute fusion of the same fragment twice leaves the scene volume unchanged (same voxel set, same TSDF)."""
    from eprecon_b200.gru_fusion import GRUFusion
    inputs, coords, _, _, _ = level2
    cfg = synth.make_cfg()
    fuse = GRUFusion(cfg, direct_substitute=True, trianing=False)
    g = torch.Generator(device="cuda").manual_seed(5)
    c = coords[torch.randperm(coords.shape[0], device="cuda", generator=g)[:100000]]
    c = c[torch.argsort((c[:, 1].long() * 96 + c[:, 2]) * 96 + c[:, 3])].long()
    tsdf = torch.rand(c.shape[0], 1, device="cuda", generator=g) * 2.4 - 1.2
    cin = {k: v for k, v in _dev(inputs).items() if k not in ("occ_list", "tsdf_list")}

    def state():
        C, F = fuse.global_volume[2]["C"].long(), fuse.global_volume[2]["F"][:, 0]
        o = torch.argsort((C[:, 0] * 100000 + C[:, 1]) * 100000 + C[:, 2])
        return C[o], F[o]
    fuse(c, tsdf, cin, 2, {}, save_mesh=False, panoptic_infos=None)
    c1, f1 = state()
    fuse(c, tsdf, cin, 2, {}, save_mesh=False, panoptic_infos=None)
    c2, f2 = state()
    assert torch.equal(c1, c2) and torch.equal(f1, f2)
    assert c1.shape[0] == int((tsdf.abs() < 1).sum())                            # exactly the active voxels of the fragment


def test_full_fragment_is_deterministic_run_to_run(cuda_lib):
    """NeuConNet.forward on the benchmark fragment twice (fresh scene each time): bit-identical sparse TSDF."""
    from eprecon_b200.neucon_network import NeuConNet
    cfg = synth.make_cfg()
    cfg.THRESHOLDS = list(synth.BENCH_THRESHOLDS)
    net = NeuConNet(cfg)
    synth.fill_parameters_(net, 1)
    net = net.cuda().train()
    inputs, fa, fb = synth.make_fragment(seed=1)
    cin, fa, fb = _dev(inputs), _dev(fa), _dev(fb)
    outs = []
    for rep in range(2):
        cin["scene"] = [f"scene_det_{rep}"]
        out, _ = net(fa, fb, cin, {})
        assert "coords" in out
        outs.append((out["coords"].clone(), out["tsdf"].clone()))
    assert outs[0][0].shape[0] > 90000
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert bool(torch.isfinite(outs[0][1]).all())
