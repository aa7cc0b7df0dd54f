"""Input side: per-view NCHW maps -> the channels-last gather buffer in one launch, zero-copy for callers that
already hold channels-last storage; the result feeds Back_Project exactly like the reference's torch.stack (neucon_network.py:364)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pack_views_matches_stack_and_permute(cuda_lib):
    from eprecon_b200 import ops
    g = torch.Generator().manual_seed(0)
    for (V, bs, C, H, W) in ((9, 1, 24, 120, 160), (9, 2, 80, 30, 40), (18, 1, 40, 45, 60), (3, 1, 5, 7, 9)):
        views = [torch.randn(bs, C, H, W, generator=g).cuda() for _ in range(V)]
        got = ops.pack_views_nhwc(views)
        want = torch.stack(views).permute(0, 1, 3, 4, 2).contiguous()
        assert got.shape == (V, bs, H, W, C) and torch.equal(got, want)


def test_channels_last_input_is_zero_copy(cuda_lib):
    from eprecon_b200 import ops
    x = torch.randn(9, 1, 30, 40, 80, device="cuda")          # [V,bs,H,W,C] storage
    as_nchw = x.permute(0, 1, 4, 2, 3)                         # what the caller sees as [V,bs,C,H,W]
    got = ops.pack_views_nhwc(as_nchw)
    assert got.data_ptr() == x.data_ptr() and torch.equal(got, x)


def test_non_contiguous_views_fall_back_to_the_transpose(cuda_lib):
    from eprecon_b200 import ops
    views = [torch.randn(1, 24, 16, 40, device="cuda")[:, :, :, ::2] for _ in range(4)]
    got = ops.pack_views_nhwc(views)
    assert torch.equal(got, torch.stack(views).permute(0, 1, 3, 4, 2).contiguous())


def test_batched_backbones_feed_neucon(cuda_lib):
    """FeatureExtractor (eprecon_b200/backbone.py): both backbones as ONE batched call each over the 9 views, per-view BatchNorm
    statistics preserved -> equal to 2 x 9 per-view calls on the GPU (cuDNN), and its lists go straight into NeuConNet.forward."""
    from eprecon_b200 import synth
    from eprecon_b200.backbone import FeatureExtractor
    from eprecon_b200.neucon_network import NeuConNet
    torch.manual_seed(0)
    fx = FeatureExtractor(alpha=1.0).cuda().train()
    inputs, fa0, fb0 = synth.make_fragment(seed=1, image_hw=(240, 320), n_vox=(64, 64, 64))
    g = torch.Generator().manual_seed(3)
    imgs = (torch.rand(1, 9, 3, 240, 320, generator=g) * 255.0).cuda()
    # fp32 convolutions on both sides (cuDNN would otherwise pick TF32 kernels, and different ones for 1 and 9 images)
    with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        fa, fb = fx(imgs)
        x = fx.normalizer(imgs)
        for v in (0, 4, 8):
            want = fx.backbone2d(x[:, v])
            for lvl in range(3):
                assert fa[v][lvl].shape == fa0[v][lvl].shape                      # the pyramid layout synth.make_fragment mimics
                scale = want[lvl].abs().max()
                assert ((fa[v][lvl] - want[lvl]).abs().max() / scale).item() < 2e-4
    cfg = synth.make_cfg(n_vox=(64, 64, 64))
    cfg.THRESHOLDS = [-100.0, -100.0, -100.0]          # random-init backbone features: keep every voxel so all levels run
    net = NeuConNet(cfg)
    synth.fill_parameters_(net, 1)
    net = net.cuda().train()
    cin = {k: (v.cuda() if torch.is_tensor(v) else ([t.cuda() for t in v] if isinstance(v, list) and torch.is_tensor(v[0]) else v))
           for k, v in inputs.items()}
    cin["scene"] = ["scene_input_side"]
    out, _ = net(fa, fb, cin, {})
    assert out is not None
