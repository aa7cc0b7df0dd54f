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
