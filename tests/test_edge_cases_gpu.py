"""Edge cases of the drop-in modules: batch size 2, degenerate fragments (the reference's None / early-return
conventions), the direct-substitute TSDF fusion of the model shell, and the small API-compat helpers."""
import numpy as np
import pytest
import torch

from eprecon_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-3


def rel(a, b):
    a, b = a.detach().cpu().float(), b.detach().cpu().float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def _dev(obj):
    if torch.is_tensor(obj):
        return obj.cuda()
    if isinstance(obj, list):
        return [_dev(o) for o in obj]
    if isinstance(obj, dict):
        return {k: _dev(v) for k, v in obj.items()}
    return obj


def test_back_project_batch_of_two(cuda_lib):
    """bs = 2: per-batch origins / KRt / feature maps, outputs concatenated in batch order like the reference loop."""
    from oracle import restate
    from eprecon_b200.occupancy_initialization import Back_Project
    inputs, fa, fb = synth.make_fragment(seed=3, image_hw=(240, 320), n_vox=(64, 64, 64), bs=2)
    g = torch.stack(torch.meshgrid(*[torch.arange(0, 64, 4)] * 3, indexing="ij")).view(3, -1)
    per = [torch.cat([torch.full((1, g.shape[1]), b, dtype=torch.long), g]).t() for b in range(2)]
    coords = torch.cat(per, 0).contiguous().int()
    feats = torch.stack([f[2] for f in fb])
    kr = inputs["proj_matrices"][:, :, 2].permute(1, 0, 2, 3).contiguous()
    origin = inputs["vol_origin_partial"]
    want = restate.backproject(coords, origin, 0.04, feats, kr, 2)
    got = Back_Project(80)(coords.cuda(), origin.cuda(), 0.04, feats.cuda(), kr.cuda(), 2)
    assert torch.equal(got[4].cpu(), want["count"]) and torch.equal(got[1].cpu(), want["coords"])
    assert torch.equal(got[3].cpu(), want["mask"])
    assert rel(got[0], want["feat"]) < RTOL
    assert set(got[1][:, 0].cpu().tolist()) == {0, 1}


def test_neucon_early_returns_like_the_reference(cuda_lib):
    from eprecon_b200.neucon_network import NeuConNet
    n_vox = (64, 64, 64)
    cfg = synth.make_cfg(n_vox=n_vox)
    net = NeuConNet(cfg)
    synth.fill_parameters_(net, 1)
    net = net.cuda()
    inputs, fa, fb = synth.make_fragment(seed=1, image_hw=(240, 320), n_vox=n_vox)
    cin, fa, fb = _dev(inputs), _dev(fa), _dev(fb)
    # (1) fragment volume nowhere near the cameras: initialisation returns None -> no 'coords', init loss key present
    far = dict(cin)
    far["vol_origin_partial"] = cin["vol_origin_partial"] + 500.0
    out, loss = net(fa, fb, far, {}, init_overlap_count=7)
    assert "coords" not in out and out["init_overlap_count"] == 7 and "occupancy_initialization_loss" in loss
    # (2) default thresholds [0,0,0] with these synthetic weights leave < 500 occupied voxels at level 0 -> early return
    out, loss = net(fa, fb, dict(cin, scene=["s_early"]), {})
    assert "coords" not in out and "tsdf_occ_loss_0" in loss
    # (3) training-only modes are refused loudly
    with pytest.raises(NotImplementedError):
        net(fa, fb, cin, {}, only_train_init=True)


def test_direct_substitute_tsdf_fusion_two_fragments(cuda_lib):
    """GRUFusion(direct_substitute=True) as models/neuralrecon.py:34,72 uses it: the scene TSDF volume after two
    overlapping fragments equals the oracle's (replace-inside-bounding-volume rule)."""
    from oracle import restate
    from eprecon_b200.gru_fusion import GRUFusion
    n_vox = (64, 64, 64)
    cfg = synth.make_cfg(n_vox=n_vox)
    fuse = GRUFusion(cfg, direct_substitute=True, trianing=False)
    state = restate.FusionState()
    g = torch.Generator().manual_seed(11)
    for frag in (0, 1):
        inputs, _, _ = synth.make_fragment(seed=1, image_hw=(240, 320), n_vox=n_vox, frag_index=frag, with_features=False)
        occ = inputs["occ_list"][0][0]
        c = torch.nonzero(occ)
        c = c[torch.randperm(len(c), generator=g)[: len(c) // 2]]
        c = c[torch.argsort((c[:, 0] * 64 + c[:, 1]) * 64 + c[:, 2])]
        coords = torch.cat([torch.zeros(len(c), 1, dtype=torch.long), c], 1)
        tsdf = (torch.rand(len(c), 1, generator=g) * 2.4 - 1.2)          # some |tsdf| >= 1 entries (inactive rows)
        want_c, want_v, _, _ = restate.gru_fusion(state, {}, cfg, coords, tsdf, inputs, 2, None, direct_substitute=True)
        outputs = fuse(coords.cuda(), tsdf.cuda(), _dev(inputs), 2, {}, save_mesh=True, panoptic_infos=None)
        gC, gF = fuse.global_volume[2]["C"].cpu().long(), fuse.global_volume[2]["F"][:, :1].cpu()
        key = lambda x: (x[:, 0] * 100000 + x[:, 1]) * 100000 + x[:, 2]  # noqa: E731
        om, ow = torch.argsort(key(gC)), torch.argsort(key(state.C[2]))
        assert torch.equal(gC[om], state.C[2][ow]) and torch.equal(gF[om], state.F[2][ow])
        vol = outputs["scene_tsdf"][-1].cpu()
        assert vol.shape == tuple((state.C[2].max(0)[0] - state.C[2].min(0)[0] + 1).tolist())


def test_api_compat_helpers(cuda_lib):
    from eprecon_b200.generate_grids import generate_grid
    from eprecon_b200.neucon_network import NeuConNet
    grid, shape = generate_grid([96, 96, 96], 2)
    assert shape == (48, 48, 48) and grid.shape == (3, 48 ** 3) and grid.dtype == torch.float32
    assert grid[:, 1].tolist() == [0.0, 0.0, 2.0] and grid[:, 48].tolist() == [0.0, 2.0, 0.0]      # x outer, z inner
    net = NeuConNet(synth.make_cfg())
    pre_c = torch.tensor([[0, 8, 16, 24], [0, 4, 4, 4]], dtype=torch.int64, device="cuda")
    pre_f = torch.arange(2 * 6, dtype=torch.float32, device="cuda").view(2, 6)
    up_f, up_c = net.upsample(pre_f, pre_c, 2)
    assert up_c.dtype == torch.int64 and up_c.shape == (16, 4) and up_f.shape == (16, 6)
    # children order of the reference: self, +x, +y, +z, +xy, +xz, +yz, +xyz
    assert up_c[:8, 1:].cpu().tolist() == [[8, 16, 24], [10, 16, 24], [8, 18, 24], [8, 16, 26], [10, 18, 24], [10, 16, 26],
                                           [8, 18, 26], [10, 18, 26]]
    assert torch.equal(up_f[:8].cpu(), pre_f[:1].cpu().expand(8, 6)) and torch.equal(up_f[8:].cpu(), pre_f[1:].cpu().expand(8, 6))
