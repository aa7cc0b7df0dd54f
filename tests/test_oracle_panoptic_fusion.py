"""Oracle pin for the panoptic global fusion: oracle/restate.gru_fusion(direct_substitute, panoptic_info)
vs tests/golden/panoptic_fusion_small.npz (the UNMODIFIED reference GRUFusion run over three overlapping fragments)."""
import os
import sys

import numpy as np
import pytest
import torch

from eprecon_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD = os.path.join(HERE, "golden", "panoptic_fusion_small.npz")


def canonical(C, *cols):
    key = (C[:, 0] * 100000 + C[:, 1]) * 100000 + C[:, 2]
    o = torch.argsort(key)
    return [C[o]] + [x.reshape(len(C), -1)[o] for x in cols]


def test_inputs_regenerate_exactly():
    from make_golden_panoptic_fusion import fusion_inputs
    g = np.load(GOLD)
    for frag in range(3):
        _, coords, tsdf, info = fusion_inputs(frag)
        assert np.array_equal(coords.numpy(), g[f"f{frag}_coords"]) and np.array_equal(tsdf.numpy(), g[f"f{frag}_tsdf"])
        assert np.array_equal(info["panoptic_seg"][0].numpy(), g[f"f{frag}_seg"])


def test_direct_substitute_panoptic_fusion_matches_reference():
    from make_golden_panoptic_fusion import N_VOX, fusion_inputs
    from oracle import restate
    g = np.load(GOLD)
    cfg = synth.make_cfg(n_vox=N_VOX)
    state = restate.FusionState()
    for frag in range(3):
        inputs, coords, tsdf, info = fusion_inputs(frag)
        inputs = {k: v for k, v in inputs.items() if k not in ("occ_list", "tsdf_list")}
        restate.gru_fusion(state, {}, cfg, coords, tsdf, inputs, 2, None, direct_substitute=True, panoptic_info=info)
        C, F, I, S = canonical(state.C[2], state.F[2], state.I, state.S)
        assert np.array_equal(C.numpy(), g[f"f{frag}_gC"]), frag
        assert np.array_equal(F.numpy(), g[f"f{frag}_gF"]), frag
        assert np.array_equal(I.numpy(), g[f"f{frag}_gI"]), frag
        assert np.array_equal(S.numpy(), g[f"f{frag}_gS"]), frag
