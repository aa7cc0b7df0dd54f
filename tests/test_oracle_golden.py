"""Pin oracle/restate.py (the travelling CPU restatement) against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py, run in the build container).  CPU only."""
import os

import numpy as np
import pytest
import torch

from eprecon_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden", "neucon_small.npz")


@pytest.fixture(scope="module")
def small_run():
    from oracle import restate
    from eprecon_b200.neucon_network import NeuConNet
    g = np.load(GOLD)
    cfg = synth.make_cfg(n_vox=tuple(int(v) for v in g["n_vox"]))
    cfg.THRESHOLDS = [float(v) for v in g["thresholds"]]
    net = NeuConNet(cfg)                    # parameter container only (no CUDA call on construction)
    sd = synth.synthetic_state_dict(net, 1)
    inputs, fa, fb = synth.make_fragment(seed=int(g["seed"]), n_views=int(g["n_views"]),
                                         image_hw=tuple(int(v) for v in g["image_hw"]),
                                         n_vox=tuple(int(v) for v in g["n_vox"]))
    trace = {}
    with torch.no_grad():
        out = restate.neucon_forward(sd, cfg, fa, fb, inputs, restate.FusionState(), trace=trace)
    return g, out, trace


def _close(g, name, x, rtol=2e-4, exact=False):
    x = x.detach()
    assert list(x.shape) == list(g[name + "_shape"]), (name, x.shape, g[name + "_shape"])
    rows = x[torch.from_numpy(g[name + "_idx"])].numpy()
    want = g[name + "_rows"]
    if exact:
        assert np.array_equal(rows.astype(np.int64), want.astype(np.int64)), name
    else:
        scale = max(1e-6, float(np.abs(want).max()))
        assert float(np.abs(rows - want).max()) / scale <= rtol, (name, float(np.abs(rows - want).max()) / scale)
        assert abs(float(x.double().abs().sum()) - float(g[name + "_abssum"])) <= rtol * float(g[name + "_abssum"]) + 1e-6


def test_init_stage_matches_reference(small_run):
    g, out, tr = small_run
    _close(g, "init_coords", tr["init"]["coords"], exact=True)
    _close(g, "init_occ", tr["init"]["occ"])
    assert np.array_equal(np.bincount(tr["init"]["count"].long().numpy(), minlength=10), g["init_count_hist"])


@pytest.mark.parametrize("level", [0, 1, 2])
def test_levels_match_reference(small_run, level):
    g, out, tr = small_run
    pre, lv = tr[f"l{level}_pre_gru"], tr[f"l{level}"]
    _close(g, f"bp{level}_coords", pre["coords"], exact=True)
    _close(g, f"spv{level}", pre["spvcnn"])
    _close(g, f"gru{level}_coords", lv["coords"], exact=True)
    _close(g, f"gru{level}_values", lv["feat_all"])
    _close(g, f"tsdf{level}", lv["tsdf"])
    _close(g, f"occ{level}", lv["occ"])


def test_final_sparse_tsdf_matches_reference(small_run):
    """Final voxel set: identical except near-ties of the last occupancy threshold (the reference's projection is an
    sgemm whose summation order is unspecified; SURVEY.md section 7 'bit-exact masks through thresholds')."""
    g, out, tr = small_run
    assert out is not None
    key = lambda c: (c[:, 1].astype(np.int64) * 4096 + c[:, 2]) * 4096 + c[:, 3]  # noqa: E731
    mine, ref = out["coords"].numpy(), g["final_coords"].astype(np.int64)
    km, kr = key(mine), key(ref)
    common, im, ir = np.intersect1d(km, kr, return_indices=True)
    n_diff = (len(km) - len(common)) + (len(kr) - len(common))
    assert n_diff <= max(4, int(1e-4 * len(kr))), n_diff
    # every voxel only one side kept must sit within 1e-4 of the threshold in the restatement
    lv = tr["l2"]
    thr = float(g["thresholds"][2])
    all_keys = key(lv["coords"].numpy())
    occ = lv["occ"].numpy().reshape(-1)
    only = np.setxor1d(km, kr)
    pos = np.searchsorted(np.sort(all_keys), only)
    margins = np.abs(occ[np.argsort(all_keys)][pos] - thr)
    assert (margins < 1e-4).all(), margins
    # order-preserving and value-exact on the common voxels
    assert np.all(np.diff(im) > 0) and np.all(np.diff(ir) > 0)
    scale = np.abs(g["final_tsdf"]).max()
    assert np.abs(out["tsdf"].numpy()[im] - g["final_tsdf"][ir]).max() <= 2e-4 * scale


def test_panoptic_feature_preparation_matches_reference(small_run):
    """Next-row work (first stage): level alignment + per-level panoptic MLPs + SubM mask features."""
    from oracle import restate
    from eprecon_b200.neucon_network import NeuConNet
    g, out, tr = small_run
    cfg = synth.make_cfg(n_vox=tuple(int(v) for v in g["n_vox"]))
    sd = synth.synthetic_state_dict(NeuConNet(cfg), 1)
    with torch.no_grad():
        pp = restate.panoptic_prepare(sd, cfg, tr)
    for p in range(3):
        # the reference's level-2 set differs from the restatement's by the 2 threshold-tie voxels (see the final-TSDF test):
        # compare shapes with that slack and the sampled rows through their coordinates' order-insensitive statistics
        want_n = int(g[f"pano{p}_shape"][0])
        assert abs(pp["feats"][p].shape[0] - want_n) <= 4 and pp["feats"][p].shape[1] == 48
        assert abs(float(pp["feats"][p].double().abs().sum()) - float(g[f"pano{p}_abssum"])) <= 2e-4 * float(g[f"pano{p}_abssum"])
    assert abs(pp["mask_features"].shape[0] - int(g["mask_feats_shape"][0])) <= 4
    assert abs(float(pp["mask_features"].double().abs().sum()) - float(g["mask_feats_abssum"])) <= 2e-4 * float(g["mask_feats_abssum"])
