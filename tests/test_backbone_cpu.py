"""Input-side shell (eprecon_b200/backbone.py): one batched backbone call over the V views of a fragment
must equal V per-view calls -- outputs AND BatchNorm buffers -- because the reference evaluates in train() mode (main.py:357)
and therefore normalises every view with its own statistics.  With /root/reference present (build container) the per-view
calls are the UNMODIFIED models/backbone.py; elsewhere this package's own per-view forward."""
import os
import sys

import pytest
import torch

from eprecon_b200.backbone import FeatureExtractor, MnasMulti, ViewBatchNorm2d

REF = "/root/reference"


def _imgs(v=3, b=1, hw=(64, 96), seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(v, b, 3, *hw, generator=g) * 255.0 - 110.0


def _randomise(net, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.3 if p.ndim > 1 else 0.1) + (1.0 if p.ndim == 1 else 0.0))
        for n, buf in net.named_buffers():
            if n.endswith("running_mean"):
                buf.copy_(torch.randn(buf.shape, generator=g) * 0.1)
            elif n.endswith("running_var"):
                buf.copy_(torch.rand(buf.shape, generator=g) + 0.5)


def _close(a, b, tol=2e-4):
    scale = b.abs().max().clamp_min(1e-6)
    assert ((a - b).abs().max() / scale).item() < tol, ((a - b).abs().max() / scale).item()


@pytest.mark.parametrize("batch", [1, 2])
@pytest.mark.parametrize("train", [True, False])
def test_batched_views_equal_per_view_calls(batch, train):
    torch.manual_seed(0)
    a, b = MnasMulti(0.5), MnasMulti(0.5)
    _randomise(a, 1)
    b.load_state_dict(a.state_dict())
    a.train(train)
    b.train(train)
    imgs = _imgs(v=3, b=batch)
    with torch.no_grad():
        per_view = [a(imgs[v]) for v in range(3)]
        batched = b.forward_views(imgs)
    for v in range(3):
        for lvl in range(3):
            assert batched[v][lvl].shape == per_view[v][lvl].shape
            _close(batched[v][lvl], per_view[v][lvl])
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:                                   # running statistics advanced by the same three updates, in view order
        if sa[k].dtype.is_floating_point:
            _close(sb[k], sa[k], 1e-5)
        else:
            assert torch.equal(sa[k], sb[k]), k
    assert all(m.views == 1 for m in b.modules() if isinstance(m, ViewBatchNorm2d))


def test_feature_extractor_layout():
    fx = FeatureExtractor(alpha=0.5)
    imgs = _imgs(v=4, b=1).transpose(0, 1).contiguous()           # [B, V, 3, H, W] as inputs['imgs']
    with torch.no_grad():
        fa, fb = fx(imgs)
    assert len(fa) == len(fb) == 4 and all(len(f) == 3 for f in fa)
    assert [tuple(t.shape) for t in fa[0]] == [(1, 16, 16, 24), (1, 24, 8, 12), (1, 40, 4, 6)]
    assert set(k.split(".")[0] for k in fx.state_dict()) == {"backbone2d", "backbone_occ_pano"}


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir(REF), reason="needs /root/reference (build container only)")
@pytest.mark.parametrize("alpha", [0.5, 1.0])
def test_matches_the_unmodified_reference_backbone(alpha, monkeypatch):
    import torchvision
    # the reference downloads ImageNet weights for alpha == 1.0; offline, hand it the bare architecture instead
    monkeypatch.setattr(torchvision.models, "mnasnet1_0", lambda **kw: torchvision.models.MNASNet(1.0))
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "models" or k.startswith("models.")}
    sys.path.insert(0, REF)
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("_ref_backbone", os.path.join(REF, "models", "backbone.py"))
        mod = importlib.util.module_from_spec(spec)
        mod.torch = torch                                          # the file uses `torch` only through these two imports
        spec.loader.exec_module(mod)
        ref = mod.MnasMulti(alpha)
    finally:
        sys.path.remove(REF)
        sys.modules.update(saved)
    mine = MnasMulti(alpha)
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    assert [tuple(v.shape) for v in ref.state_dict().values()] == [tuple(v.shape) for v in mine.state_dict().values()]
    _randomise(ref, 3)
    mine.load_state_dict(ref.state_dict())
    ref.train()
    mine.train()
    imgs = _imgs(v=3, b=1, seed=5)
    with torch.no_grad():
        want = [ref(imgs[v]) for v in range(3)]
        got = mine.forward_views(imgs)
    for v in range(3):
        for lvl in range(3):
            _close(got[v][lvl], want[v][lvl])
    for k, t in ref.state_dict().items():
        if t.dtype.is_floating_point:
            _close(mine.state_dict()[k], t, 1e-5)
