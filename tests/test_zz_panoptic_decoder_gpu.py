"""Panoptic decoder on CUDA vs the reference fixture (tests/golden/mask3dformer_small.npz, outputs of
the UNMODIFIED models/mask3dformer.py) and vs the CPU oracle at the small golden NeuConNet configuration."""
import os

import numpy as np
import pytest
import torch

from eprecon_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-3
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "mask3dformer_small.npz")
GOLD_NET = os.path.join(HERE, "golden", "neucon_small.npz")


def rel(a, b):
    a, b = torch.as_tensor(a).detach().cpu().float(), torch.as_tensor(b).detach().cpu().float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def c4(xyz):
    return torch.cat([torch.zeros_like(xyz[:, :1]), xyz], 1).to(torch.int32).contiguous().cuda()


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def decoder(cuda_lib):
    from eprecon_b200.mask3dformer import MultiScaleMaskedTransformerDecoder
    dec = MultiScaleMaskedTransformerDecoder(mask_classification=True, num_classes=20, hidden_dim=48, num_queries=80, nheads=8,
                                             dim_feedforward=192, dec_layers=6, pre_norm=False, mask_dim=48)
    synth.fill_parameters_(dec, 1, prefix="panoptic.")
    return dec.cuda()


def test_nearest_fine_index_hash_join(cuda_lib, gold):
    from oracle import restate
    from eprecon_b200.mask3dformer import nearest_fine_index
    coords = [torch.from_numpy(gold[f"coords{l}"].astype(np.int64)) for l in range(3)]
    i0 = nearest_fine_index(c4(coords[0]), c4(coords[2]), 5)
    i1 = nearest_fine_index(c4(coords[1]), c4(coords[2]), 1)
    assert np.array_equal(i0.cpu().numpy(), gold["index0"])        # the reference's cdist + argmin, bit-exact
    assert np.array_equal(i1.cpu().numpy(), gold["index1"])
    # random sparse set with many equidistant candidates: first-row tie rule
    g = torch.Generator().manual_seed(5)
    fine = torch.unique(torch.randint(0, 40, (6000, 3), generator=g), dim=0)
    fine = fine[torch.randperm(len(fine), generator=g)]
    for step, radius in ((2, 1), (4, 5)):
        coarse = torch.unique(torch.floor_divide(fine, step) * step, dim=0)
        coarse = coarse[torch.randperm(len(coarse), generator=g)]
        want = restate.nearest_fine_index(coarse, fine)
        got = nearest_fine_index(c4(coarse), c4(fine), radius)
        assert torch.equal(got.cpu(), want), step


def test_decoder_matches_reference_fixture(decoder, gold):
    coords = [torch.from_numpy(gold[f"coords{l}"].astype(np.int64)).cuda() for l in range(3)]
    feats = [torch.from_numpy(gold[f"feats{l}"]).cuda() for l in range(3)]
    mf = torch.from_numpy(gold["mask_features"]).cuda()
    dim = int(gold["dim"])
    out = decoder([f.unsqueeze(0).permute(0, 2, 1) for f in feats], [c.unsqueeze(0) for c in coords],
                  mf.unsqueeze(0).permute(0, 2, 1), (dim, dim, dim))
    assert rel(out["pred_logits"][0], gold["pred_logits"]) < RTOL
    assert rel(out["pred_masks"][0], gold["pred_masks"]) < RTOL
    assert len(out["aux_outputs"]) == 6
    for j, a in enumerate(out["aux_outputs"]):
        assert rel(a["pred_logits"][0], gold["aux_logits"][j]) < RTOL, j
    from eprecon_b200.mask3dformer import panoptic_post
    seg, info = panoptic_post(out)["panoptic_seg"]
    assert seg.dtype == torch.int32 and (seg.cpu().numpy() != gold["post_seg"]).mean() <= 1e-3


def _decoder_inputs(gold):
    coords = [torch.from_numpy(gold[f"coords{l}"].astype(np.int64)).cuda() for l in range(3)]
    feats = [torch.from_numpy(gold[f"feats{l}"]).cuda() for l in range(3)]
    mf = torch.from_numpy(gold["mask_features"]).cuda()
    dim = int(gold["dim"])
    return ([f.unsqueeze(0).permute(0, 2, 1) for f in feats], [c.unsqueeze(0) for c in coords], mf.unsqueeze(0).permute(0, 2, 1),
            (dim, dim, dim))


def test_native_decoder_aux_masks_match_reference_fixture(decoder, gold):
    """The six intermediate mask predictions of the one-call decoder (csrc/decoder.cu) against the abs-sums the UNMODIFIED
    reference produced, and the final prediction without the aux outputs requested (the masks buffer is then not written)."""
    from eprecon_b200 import mask3dformer
    assert mask3dformer.NATIVE_DECODER
    out = decoder(*_decoder_inputs(gold))
    for j, a in enumerate(out["aux_outputs"]):
        got = a["pred_masks"][0].abs().sum().item()
        assert abs(got - float(gold["aux_masks_abssum"][j])) <= RTOL * float(gold["aux_masks_abssum"][j]), j
    decoder.aux_outputs = False
    try:
        lean = decoder(*_decoder_inputs(gold))
    finally:
        decoder.aux_outputs = True
    assert "aux_outputs" not in lean
    assert torch.equal(lean["pred_masks"], out["pred_masks"]) and torch.equal(lean["pred_logits"], out["pred_logits"])


def test_native_decoder_matches_per_op_formulation(decoder, gold, monkeypatch):
    """One native call vs the per-op route (fused cross-attention + ATen linears / LayerNorm / softmax): every prediction,
    incl. a case where every query sees either all keys blocked or none (rank-one mask features: the sign of a query's mask
    logit is then the same for every voxel) -> the fully blocked ones attend everywhere (mask3dformer.py:392)."""
    from eprecon_b200 import mask3dformer
    args = list(_decoder_inputs(gold))
    for variant in ("fixture", "blocked"):
        if variant == "blocked":
            g = torch.Generator().manual_seed(3)
            scale = (0.5 + torch.rand(args[2].shape[2], generator=g)).cuda()
            args[2] = (torch.ones_like(args[2]) * scale.view(1, 1, -1)).contiguous()
        native = decoder(*args)
        monkeypatch.setattr(mask3dformer, "NATIVE_DECODER", False)
        per_op = decoder(*args)
        monkeypatch.setattr(mask3dformer, "NATIVE_DECODER", True)
        assert rel(native["pred_logits"], per_op["pred_logits"]) < RTOL, variant
        assert rel(native["pred_masks"], per_op["pred_masks"]) < RTOL, variant
        if variant == "blocked":
            sign = per_op["aux_outputs"][0]["pred_masks"][0] < 0
            assert bool((sign.all(1) | (~sign).all(1)).all()) and bool(sign.all(1).any()) and bool((~sign).all(1).any())
        for j in range(6):
            assert rel(native["aux_outputs"][j]["pred_logits"], per_op["aux_outputs"][j]["pred_logits"]) < RTOL, (variant, j)
            assert rel(native["aux_outputs"][j]["pred_masks"], per_op["aux_outputs"][j]["pred_masks"]) < RTOL, (variant, j)


def test_native_decoder_rejects_other_shapes(cuda_lib, gold):
    from eprecon_b200._lib import EpreconError
    from eprecon_b200.mask3dformer import MultiScaleMaskedTransformerDecoder
    dec = MultiScaleMaskedTransformerDecoder(mask_classification=True, num_classes=20, hidden_dim=48, num_queries=80, nheads=8,
                                             dim_feedforward=192, dec_layers=3, pre_norm=False, mask_dim=48).cuda()
    with pytest.raises(EpreconError):
        dec(*_decoder_inputs(gold))


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_panoptic_inference_matches_reference(cuda_lib, gold, case):
    from eprecon_b200.mask3dformer import panoptic_inference
    seg, info = panoptic_inference(torch.from_numpy(gold[f"pi{case}_cls"]).cuda(), torch.from_numpy(gold[f"pi{case}_msk"]).cuda())
    assert np.array_equal(seg.cpu().numpy(), gold[f"pi{case}_seg"])
    got = np.asarray([[d["id"], int(d["isthing"]), d["category_id"]] for d in info], dtype=np.int32).reshape(-1, 3)
    assert np.array_equal(got, gold[f"pi{case}_info"])


def test_decoder_rejects_cpu_tensors():
    from eprecon_b200._lib import EpreconError
    from eprecon_b200.mask3dformer import MultiScaleMaskedTransformerDecoder
    dec = MultiScaleMaskedTransformerDecoder(mask_classification=True, num_classes=20, hidden_dim=48, num_queries=80, nheads=8,
                                             dim_feedforward=192, dec_layers=1, pre_norm=False, mask_dim=48)
    x = torch.zeros(1, 48, 10)
    with pytest.raises(EpreconError):
        dec([x, x, x], [torch.zeros(1, 10, 3, dtype=torch.long)] * 3, x, (8, 8, 8))


def test_neucon_forward_with_panoptic_head(cuda_lib):
    """NeuConNet.forward with the full panoptic head on the small golden configuration: the decoder runs on the product's
    own level-aligned features (checked against the oracle in test_neucon_gpu.py) and must agree with the oracle decoder
    fed the same features."""
    from oracle import restate
    from eprecon_b200.neucon_network import NeuConNet
    g = np.load(GOLD_NET)
    n_vox = tuple(int(v) for v in g["n_vox"])
    cfg = synth.make_cfg(n_vox=n_vox)
    cfg.THRESHOLDS = [float(v) for v in g["thresholds"]]
    net = NeuConNet(cfg)
    sd = synth.synthetic_state_dict(net, 1)
    net = net.cuda()
    net.with_panoptic = True
    inputs, fa, fb = synth.make_fragment(seed=int(g["seed"]), n_views=int(g["n_views"]),
                                         image_hw=tuple(int(v) for v in g["image_hw"]), n_vox=n_vox)
    cin = {k: (v.cuda() if torch.is_tensor(v) else ([t.cuda() for t in v] if isinstance(v, list) and torch.is_tensor(v[0]) else v))
           for k, v in inputs.items()}
    cin["scene"] = ["scene_pano_full"]
    out, _ = net([[t.cuda() for t in f] for f in fa], [[t.cuda() for t in f] for f in fb], cin, {})
    assert "coords" in out and isinstance(out["panoptic_info"], list) and len(out["panoptic_info"]) == 1
    seg, info = out["panoptic_info"][0]["panoptic_seg"]
    pf = out["panoptic_features"]
    assert seg.shape[0] == pf["coords"][2].shape[0] == out["coords"].shape[0]
    feats = [pf["feats"][p][:, :48].cpu() for p in range(3)]
    xyz = [pf["coords"][p][:, 1:].cpu() for p in range(3)]
    with torch.no_grad():
        want = restate.mask3dformer(sd, "panoptic", feats, xyz, pf["mask_features"][:, :48].cpu(), n_vox)
    got = out["panoptic_outs"][0]
    assert rel(got["pred_logits"][0], want["pred_logits"]) < RTOL
    assert rel(got["pred_masks"][0], want["pred_masks"]) < RTOL
    wseg, winfo = restate.panoptic_inference(want["pred_logits"], want["pred_masks"])
    assert (seg.cpu() != wseg).float().mean().item() <= 1e-3
    assert [d["category_id"] for d in info] == [d["category_id"] for d in winfo]


def test_direct_substitute_panoptic_fusion_matches_reference(cuda_lib):
    """GRUFusion(direct_substitute=True) with panoptic_infos over three overlapping fragments -- scene
    TSDF, instance and semantic ids equal the UNMODIFIED reference's (tests/golden/panoptic_fusion_small.npz), bit-exact."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_golden_panoptic_fusion import N_VOX, fusion_inputs
    from eprecon_b200.gru_fusion import GRUFusion
    g = np.load(os.path.join(HERE, "golden", "panoptic_fusion_small.npz"))
    cfg = synth.make_cfg(n_vox=N_VOX)
    fuse = GRUFusion(cfg, direct_substitute=True, trianing=False)
    outputs = {}
    for frag in range(3):
        inputs, coords, tsdf, info = fusion_inputs(frag)
        cin = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in inputs.items() if k not in ("occ_list", "tsdf_list")}
        pinfo = {"panoptic_seg": [info["panoptic_seg"][0].cuda(), info["panoptic_seg"][1]]}
        outputs = fuse(coords.cuda(), tsdf.cuda(), cin, 2, outputs, save_mesh=True, panoptic_infos=[pinfo])
        C = fuse.global_volume[2]["C"].cpu().long()
        key = (C[:, 0] * 100000 + C[:, 1]) * 100000 + C[:, 2]
        o = torch.argsort(key)
        assert np.array_equal(C[o].numpy(), g[f"f{frag}_gC"]), frag
        assert np.array_equal(fuse.global_volume[2]["F"][:, :1].cpu()[o].numpy(), g[f"f{frag}_gF"]), frag
        assert np.array_equal(fuse.global_instance.cpu()[o].numpy().reshape(-1, 1), g[f"f{frag}_gI"]), frag
        assert np.array_equal(fuse.global_semantic.cpu()[o].numpy().reshape(-1, 1), g[f"f{frag}_gS"]), frag
        assert tuple(outputs["scene_tsdf"][-1].shape) == tuple(int(v) for v in g[f"f{frag}_scene_shape"])
        assert np.array_equal(np.bincount(outputs["scene_instance"][-1].cpu().long().flatten().numpy()), g[f"f{frag}_scene_instance_hist"])
        assert np.array_equal(np.bincount(outputs["scene_semantic"][-1].cpu().long().flatten().numpy()), g[f"f{frag}_scene_semantic_hist"])


def test_direct_substitute_panoptic_fusion_batch_of_two(cuda_lib):
    """bs = 2: fragments 0 and 1 of one scene as ONE batched call.  The reference loops over the batch entries against the
    same global state (models/gru_fusion.py:274-369) and pairs panoptic_infos[i]['panoptic_seg'][0] row for row with entry
    i's voxels (:355), so the state after the call equals the fixture's state after fragment 1."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_golden_panoptic_fusion import N_VOX, fusion_inputs
    from eprecon_b200.gru_fusion import GRUFusion
    g = np.load(os.path.join(HERE, "golden", "panoptic_fusion_small.npz"))
    cfg = synth.make_cfg(n_vox=N_VOX)
    fuse = GRUFusion(cfg, direct_substitute=True, trianing=False)
    parts = [fusion_inputs(f) for f in (0, 1)]
    cin = {}
    for k in ("proj_matrices", "vol_origin_partial", "vol_origin", "world_to_aligned_camera"):
        cin[k] = torch.cat([p[0][k] for p in parts]).cuda()
    cin["scene"] = [p[0]["scene"][0] for p in parts]
    cin["fragment"] = [p[0]["fragment"][0] for p in parts]
    coords = torch.cat([torch.cat([torch.full((len(p[1]), 1), b, dtype=torch.long), p[1][:, 1:]], 1) for b, p in enumerate(parts)])
    tsdf = torch.cat([p[2] for p in parts])
    infos = [{"panoptic_seg": [p[3]["panoptic_seg"][0].cuda(), p[3]["panoptic_seg"][1]]} for p in parts]
    fuse(coords.cuda(), tsdf.cuda(), cin, 2, {}, save_mesh=False, panoptic_infos=infos)
    C = fuse.global_volume[2]["C"].cpu().long()
    o = torch.argsort((C[:, 0] * 100000 + C[:, 1]) * 100000 + C[:, 2])
    assert np.array_equal(C[o].numpy(), g["f1_gC"])
    assert np.array_equal(fuse.global_volume[2]["F"][:, :1].cpu()[o].numpy(), g["f1_gF"])
    assert np.array_equal(fuse.global_instance.cpu()[o].numpy().reshape(-1, 1), g["f1_gI"])
    assert np.array_equal(fuse.global_semantic.cpu()[o].numpy().reshape(-1, 1), g["f1_gS"])


@pytest.mark.parametrize("n_keys,n_queries,n_heads", [(1, 80, 8), (63, 80, 8), (65, 96, 8), (5000, 80, 8), (105001, 80, 8), (777, 7, 2)])
def test_fused_masked_attention_matches_fp64_softmax(cuda_lib, n_keys, n_queries, n_heads):
    """csrc/attention.cu vs a float64 masked softmax-attention (the nn.MultiheadAttention math of mask3dformer.py:70-130)."""
    from eprecon_b200 import ops
    g = torch.Generator().manual_seed(n_keys)
    E = n_heads * 6
    q = torch.randn(n_queries, E, generator=g) * 2.0
    k = torch.randn(n_keys, E, generator=g) * 2.0
    v = torch.randn(n_keys, E, generator=g)
    blocked = torch.rand(n_queries, n_keys, generator=g) < 0.6
    blocked[:, 0] = False                                   # every query sees at least one key
    if n_keys > 200:
        blocked[3, 1:] = True                               # a query with a single visible key
        blocked[5, 64:192] = True                           # whole tiles masked
        blocked[:, 128:136] = True                          # a softmax block masked for every query
    scale = 1.0 / 6 ** 0.5
    qh = (q.double() * scale).view(n_queries, n_heads, 6).transpose(0, 1)
    kh = k.double().view(n_keys, n_heads, 6).transpose(0, 1)
    vh = v.double().view(n_keys, n_heads, 6).transpose(0, 1)
    s = (qh @ kh.transpose(1, 2)).masked_fill(blocked.unsqueeze(0), float("-inf"))
    want = (torch.softmax(s, -1) @ vh).transpose(0, 1).reshape(n_queries, E)
    got = ops.masked_attention(q.cuda(), k.cuda(), v.cuda(), blocked.cuda(), n_heads, scale).cpu().double()
    assert (got - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())
    got2 = ops.masked_attention(q.cuda(), k.cuda(), v.cuda(), None, n_heads, scale).cpu().double()
    want2 = (torch.softmax(qh @ kh.transpose(1, 2), -1) @ vh).transpose(0, 1).reshape(n_queries, E)
    assert (got2 - want2).abs().max().item() <= 2e-5 * max(1.0, want2.abs().max().item())
