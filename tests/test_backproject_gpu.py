"""CUDA back-projection vs the CPU oracle (bit-exact indices / masks / counts; floats <= 1e-3 rel)."""
import pytest
import torch

from eprecon_b200 import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-3  # north-star tolerance for feature floats


def _level_inputs(level, seed=1, image_hw=(480, 640), n_vox=(96, 96, 96), stride=None, n_views=9):
    inputs, fa, fb = synth.make_fragment(seed=seed, image_hw=image_hw, n_vox=n_vox, n_views=n_views)
    scale = 2 - level
    interval = stride or 2 ** scale
    axes = [torch.arange(0, n_vox[a], interval) for a in range(3)]
    g = torch.stack(torch.meshgrid(*axes, indexing="ij")).view(3, -1)
    coords = torch.cat([torch.zeros(1, g.shape[1], dtype=torch.long), g]).t().contiguous().int()
    feats = torch.stack([f[scale] for f in fb])
    kr = inputs["proj_matrices"][:, :, scale].permute(1, 0, 2, 3).contiguous()
    return coords, inputs["vol_origin_partial"], feats, kr


def _rel_err(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


@pytest.mark.parametrize("level,min_views", [(0, 2), (1, 0), (2, 0)])
def test_back_project_matches_oracle(cuda_lib, level, min_views):
    from oracle import restate
    from eprecon_b200.occupancy_initialization import Back_Project
    stride = {0: 4, 1: 2, 2: 2}[level]  # level 2 at stride 2 keeps the oracle at a few seconds
    coords, origin, feats, kr = _level_inputs(level, stride=stride)
    want = restate.backproject(coords, origin, 0.04, feats, kr, min_views)
    got = Back_Project(feats.shape[2])(coords.cuda(), origin.cuda(), 0.04, feats.cuda(), kr.cuda(), min_views)
    feat, oc, im_grid, mask, count = got
    assert torch.equal(count.cpu(), want["count"])                      # bit-exact view counts
    assert torch.equal(oc.cpu(), want["coords"])                        # bit-exact surviving indices, same order
    assert torch.equal(mask.cpu(), want["mask"])                        # bit-exact visibility masks
    vis = want["mask"]
    assert torch.equal(im_grid.cpu()[vis], want["im_grid"][vis])        # identical sample positions where used
    assert _rel_err(feat.cpu(), want["feat"]) <= RTOL


def test_init_stage_variance_matches_oracle(cuda_lib):
    from oracle import restate
    from eprecon_b200 import ops
    coords, origin, _, kr = _level_inputs(1, stride=2)
    feats = torch.randn(9, 1, 32, 60, 80, generator=torch.Generator().manual_seed(3))
    want = restate.backproject(coords, origin, 0.04, feats, kr, 2, mode="meanvar")
    got = ops.backproject(coords.cuda(), origin.cuda(), 0.04, ops.to_nhwc(feats.cuda()), kr.cuda(), 2, mode="meanvar")
    assert torch.equal(got["coords"].cpu(), want["coords"])
    assert _rel_err(got["feat"].cpu(), want["feat"]) <= RTOL


def test_legacy_back_project_depth_channel(cuda_lib):
    from oracle import restate
    from eprecon_b200.back_project import back_project
    coords, origin, feats, kr = _level_inputs(0, stride=4)
    want = restate.backproject(coords, origin, 0.04, feats, kr, 2)
    zn = restate.legacy_depth_channel(want["zbar"])
    feat, oc, count = back_project(coords.cuda(), origin.cuda(), 0.04, feats.cuda(), kr.cuda(), 2)
    assert oc.dtype == torch.float32 and torch.equal(oc.cpu(), want["coords"].float())
    assert torch.equal(count.cpu(), want["count"])
    assert _rel_err(feat[:, :-1].cpu(), want["feat"]) <= RTOL
    assert (feat[:, -1:].cpu() - zn).abs().max().item() <= 1e-5


def test_degenerate_fragment_returns_none(cuda_lib):
    from eprecon_b200.occupancy_initialization import Back_Project
    coords, origin, feats, kr = _level_inputs(0, stride=4)
    far = origin + 1000.0  # volume nowhere near the cameras: no voxel visible
    assert Back_Project(80)(coords.cuda(), far.cuda(), 0.04, feats.cuda(), kr.cuda(), 2) is None


def test_generic_channel_count_and_18_views(cuda_lib):
    """Non-specialised C (generic kernel) and the 18-view stress shape (config 5)."""
    from oracle import restate
    from eprecon_b200 import ops
    inputs, _, _ = synth.make_fragment(seed=2, n_views=18, image_hw=(720, 960), n_vox=(128, 128, 128),
                                       with_features=False)
    g = torch.stack(torch.meshgrid(*[torch.arange(0, 128, 4)] * 3, indexing="ij")).view(3, -1)
    coords = torch.cat([torch.zeros(1, g.shape[1], dtype=torch.long), g]).t().contiguous().int()
    feats = torch.randn(18, 1, 12, 45, 60, generator=torch.Generator().manual_seed(5))
    kr = inputs["proj_matrices"][:, :, 2].permute(1, 0, 2, 3).contiguous()
    origin = inputs["vol_origin_partial"]
    want = restate.backproject(coords, origin, 0.04, feats, kr, 2)
    got = ops.backproject(coords.cuda(), origin.cuda(), 0.04, ops.to_nhwc(feats.cuda()), kr.cuda(), 2)
    assert torch.equal(got["count"].cpu(), want["count"])
    assert torch.equal(got["coords"].cpu(), want["coords"])
    assert _rel_err(got["feat"].cpu(), want["feat"]) <= RTOL
