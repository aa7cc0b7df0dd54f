"""Native executor (csrc/executor.cu: one C call per reference module) vs the per-kernel Python programs in
eprecon_b200/modules.py: the same launchers in the same order => every stage must be BIT-identical."""
import os

import numpy as np
import pytest
import torch

from eprecon_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "neucon_small.npz")


def _cuda(obj):
    if torch.is_tensor(obj):
        return obj.cuda()
    if isinstance(obj, list):
        return [_cuda(o) for o in obj]
    if isinstance(obj, dict):
        return {k: _cuda(v) for k, v in obj.items()}
    return obj


def _run(net, fa, fb, ins, scene, use_exec):
    from eprecon_b200 import executor
    old = executor.ENABLED
    executor.ENABLED = use_exec
    try:
        net.trace, net.teacher = {}, None
        out, _ = net(fa, fb, dict(ins, scene=[scene]), {})
        torch.cuda.synchronize()
        return out, net.trace
    finally:
        executor.ENABLED = old
        net.trace = None


def _same(a, b, path=""):
    if torch.is_tensor(a):
        assert a.shape == b.shape, path
        assert torch.equal(a, b), f"{path}: max abs diff {(a.float() - b.float()).abs().max().item():.3e}"
    elif isinstance(a, dict):
        assert a.keys() == b.keys(), path
        for k in a:
            _same(a[k], b[k], f"{path}/{k}")
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _same(x, y, f"{path}[{i}]")
    else:
        assert a is None and b is None or a == b, path


@pytest.mark.parametrize("with_pano", [False, True])
def test_executor_bit_identical_to_python_programs(cuda_lib, with_pano):
    from eprecon_b200 import executor
    from eprecon_b200.neucon_network import NeuConNet
    g = np.load(GOLD)
    n_vox = tuple(int(v) for v in g["n_vox"])
    cfg = synth.make_cfg(n_vox=n_vox)
    cfg.THRESHOLDS = [float(v) for v in g["thresholds"]]
    net = NeuConNet(cfg)
    synth.fill_parameters_(net, 1)
    net = net.cuda().train()
    net.with_panoptic_features = with_pano
    inputs, fa, fb = synth.make_fragment(seed=int(g["seed"]), n_views=int(g["n_views"]),
                                         image_hw=tuple(int(v) for v in g["image_hw"]), n_vox=n_vox)
    fa, fb, ins = _cuda(fa), _cuda(fb), _cuda(inputs)
    # second fragment of the SAME scene: goes through the recurrent GRU state (global rows present in the union)
    inputs2, fa2, fb2 = synth.make_fragment(seed=int(g["seed"]), n_views=int(g["n_views"]), frag_index=1,
                                            image_hw=tuple(int(v) for v in g["image_hw"]), n_vox=n_vox)
    fa2, fb2, ins2 = _cuda(fa2), _cuda(fb2), _cuda(inputs2)
    out_p, tr_p = _run(net, fa, fb, ins, "py", False)
    out_p2, tr_p2 = _run(net, fa2, fb2, ins2, "py", False)
    out_e, tr_e = _run(net, fa, fb, ins, "ex", True)
    out_e2, tr_e2 = _run(net, fa2, fb2, ins2, "ex", True)
    assert "coords" in out_p and "coords" in out_e
    assert executor.arena_peak_bytes() > 0                      # the native path really ran
    _same(tr_e, tr_p, "trace")
    _same(out_e["coords"], out_p["coords"], "coords")
    _same(out_e["tsdf"], out_p["tsdf"], "tsdf")
    if with_pano:
        _same(out_e["panoptic_features"], out_p["panoptic_features"], "panoptic_features")
    _same(tr_e2, tr_p2, "trace2")
    assert ("coords" in out_e2) == ("coords" in out_p2)
    if "coords" in out_p2:
        _same(out_e2["coords"], out_p2["coords"], "coords2")
        _same(out_e2["tsdf"], out_p2["tsdf"], "tsdf2")


@pytest.mark.parametrize("shift,batches,want_tier", [(0.0, 1, 0), (0.0, 2, 1), (8.0, 1, 1), (30.0, 1, 2)])
def test_executor_sort_key_tiers_keep_the_row_order(cuda_lib, shift, batches, want_tier):
    """The executor sorts voxels by compact Z-order keys (24 / 32 bits) when the cloud fits and falls back to wider keys
    when it does not; every tier must give the rows in the order of the 64-bit keys the Python programs use -> the module
    output is bit-identical whatever tier ends up in force."""
    from eprecon_b200 import executor
    from eprecon_b200.modules import SPVCNN
    from eprecon_b200.tensor import PointTensor
    torch.manual_seed(0)
    net = SPVCNN(num_classes=1, in_channels=24, pres=1, cr=0.25, vres=0.04, dropout=False).cuda()
    n = 20000
    g = torch.Generator().manual_seed(4)
    xyz = (torch.rand(n, 3, generator=g) - 0.5) * 4.0 + shift              # +-2 m = +-50 voxels around `shift` metres
    b = torch.randint(0, batches, (n, 1), generator=g).float()
    pts = torch.cat([xyz, b], 1).cuda()
    feat = torch.randn(n, 24, generator=g).cuda()
    old = executor.ENABLED
    L = cuda_lib
    L.ep_exec_key_tier(0)
    try:
        executor.ENABLED = True
        got = net(PointTensor(feat, pts)).clone()
        assert L.ep_exec_key_tier(-1) == want_tier
        again = net(PointTensor(feat, pts)).clone()                        # starts in the raised tier: no retry
        executor.ENABLED = False
        want = net(PointTensor(feat, pts)).clone()
    finally:
        executor.ENABLED = old
        L.ep_exec_key_tier(0)
    assert torch.equal(got, want) and torch.equal(again, want)


def test_executor_grows_its_arena(cuda_lib):
    """A scratch arena that is too small is reported (EP_ERR_WORKSPACE) and grown, never overrun."""
    from eprecon_b200 import executor
    from eprecon_b200.modules import SPVCNN
    from eprecon_b200.tensor import PointTensor
    torch.manual_seed(0)
    net = SPVCNN(num_classes=1, in_channels=24, pres=1, cr=0.25, vres=0.04, dropout=False).cuda()
    n = 20000
    pts = torch.cat([torch.rand(n, 3) * 2.0, torch.zeros(n, 1)], 1).cuda()
    feat = torch.randn(n, 24).cuda()
    old_mb, old_en = executor.ARENA_MB, executor.ENABLED
    st = executor._state()
    saved = dict(st["arenas"])
    try:
        executor.ENABLED = True
        st["arenas"].clear()
        executor.ARENA_MB = 1                                   # 1 MiB: far too small for 20 k points
        small = net(PointTensor(feat, pts)).clone()
        grown = max(a.numel() for a in st["arenas"].values())
        assert grown > (1 << 20)
        st["arenas"].clear()
        executor.ARENA_MB = 256
        big = net(PointTensor(feat, pts)).clone()
        assert torch.equal(small, big)
        executor.ENABLED = False
        ref = net(PointTensor(feat, pts))
        assert torch.equal(big, ref)
    finally:
        executor.ARENA_MB, executor.ENABLED = old_mb, old_en
        st["arenas"].clear()
        st["arenas"].update(saved)
