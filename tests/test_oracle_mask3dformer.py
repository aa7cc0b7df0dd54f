"""Oracle pin for the panoptic decoder: oracle/restate.py vs tests/golden/mask3dformer_small.npz, which
holds outputs of the UNMODIFIED reference modules (models/mask3dformer.py, plain ATen -> pinned directly)."""
import os

import numpy as np
import pytest
import torch

from eprecon_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mask3dformer_small.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def decoder_state_dict():
    """Seed-1 weights under the reference's state-dict names (prefix 'panoptic.'), from the product module."""
    from eprecon_b200.mask3dformer import MultiScaleMaskedTransformerDecoder
    dec = MultiScaleMaskedTransformerDecoder(mask_classification=True, num_classes=20, hidden_dim=48, num_queries=80, nheads=8,
                                             dim_feedforward=192, dec_layers=6, pre_norm=False, mask_dim=48)
    synth.fill_parameters_(dec, 1, prefix="panoptic.")
    return {"panoptic." + k: v.detach().clone() for k, v in dec.state_dict().items()}


def gold_inputs(g):
    coords = [torch.from_numpy(g[f"coords{l}"].astype(np.int64)) for l in range(3)]
    feats = [torch.from_numpy(g[f"feats{l}"]) for l in range(3)]
    return coords, feats, torch.from_numpy(g["mask_features"])


def test_nearest_fine_index_matches_reference_cdist_argmin(gold):
    from oracle import restate
    coords, _, _ = gold_inputs(gold)
    assert np.array_equal(restate.nearest_fine_index(coords[0], coords[2]).numpy(), gold["index0"])
    assert np.array_equal(restate.nearest_fine_index(coords[1], coords[2]).numpy(), gold["index1"])


def test_decoder_matches_reference(gold):
    from oracle import restate
    coords, feats, mf = gold_inputs(gold)
    sd = decoder_state_dict()
    dim = int(gold["dim"])
    with torch.no_grad():
        out = restate.mask3dformer(sd, "panoptic", feats, coords, mf, (dim, dim, dim))
    scale_l, scale_m = np.abs(gold["pred_logits"]).max(), np.abs(gold["pred_masks"]).max()
    assert np.abs(out["pred_logits"].numpy() - gold["pred_logits"]).max() <= 1e-4 * scale_l
    assert np.abs(out["pred_masks"].numpy() - gold["pred_masks"]).max() <= 1e-4 * scale_m
    for j, (lg, mk) in enumerate(out["aux"]):
        assert np.abs(lg.numpy() - gold["aux_logits"][j]).max() <= 1e-4 * scale_l, j
        assert abs(float(mk.double().abs().sum()) - float(gold["aux_masks_abssum"][j])) <= 1e-4 * float(gold["aux_masks_abssum"][j]), j
    seg, _ = restate.panoptic_inference(out["pred_logits"], out["pred_masks"])
    assert (seg.numpy() != gold["post_seg"]).mean() <= 1e-3      # argmax / 0.5 ties under 1e-4 float differences


@pytest.mark.parametrize("case", [0, 1, 2, 3])
def test_panoptic_inference_matches_reference(gold, case):
    from oracle import restate
    seg, info = restate.panoptic_inference(torch.from_numpy(gold[f"pi{case}_cls"]), torch.from_numpy(gold[f"pi{case}_msk"]))
    assert np.array_equal(seg.numpy(), gold[f"pi{case}_seg"])
    got = np.asarray([[d["id"], int(d["isthing"]), d["category_id"]] for d in info], dtype=np.int32).reshape(-1, 3)
    assert np.array_equal(got, gold[f"pi{case}_info"])
