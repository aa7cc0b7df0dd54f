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


def test_product_pair_statistics_host_logic_matches_oracle():
    """Host logic of eprecon_b200.gru_fusion.GRUFusion.panoptic_fusion (bincount tables + the sequential matching loop)
    against the oracle's literal restatement, on the union / row maps the oracle computes for the golden fragments.  The
    CUDA union kernel that produces `row_b` on the product path is covered by the GPU tests."""
    from make_golden_panoptic_fusion import N_VOX, fusion_inputs
    from oracle import restate
    from eprecon_b200.gru_fusion import GRUFusion
    cfg = synth.make_cfg(n_vox=N_VOX)
    fuse = GRUFusion(cfg, direct_substitute=True, trianing=False)
    state = restate.FusionState()
    dims = torch.tensor(N_VOX)
    for frag in range(3):
        inputs, coords, tsdf, info = fusion_inputs(frag)
        inputs = {k: v for k, v in inputs.items() if k not in ("occ_list", "tsdf_list")}
        # state BEFORE this fragment -> product-side tables
        fuse.global_instance, fuse.global_semantic = state.I.clone(), state.S.clone()
        gC_before = state.C[2].clone() if state.C[2] is not None else torch.zeros(0, 3, dtype=torch.long)
        restate.gru_fusion(state, {}, cfg, coords, tsdf, inputs, 2, None, direct_substitute=True,
                           panoptic_info={"panoptic_seg": [info["panoptic_seg"][0].clone(), info["panoptic_seg"][1]]})
        n_new = len(state.C[2]) - 0
        rel = ((inputs["vol_origin_partial"][0] - inputs["vol_origin"][0]) / 0.04).long()
        # the union sites are the rows the oracle appended last (in raster order)
        gl = gC_before - rel
        valid = ((gl < dims) & (gl >= 0)).all(-1)
        u = len(state.C[2]) - int((~valid).sum())
        upd = state.C[2][-u:] - rel
        lin = lambda x: (x[:, 0] * N_VOX[1] + x[:, 1]) * N_VOX[2] + x[:, 2]  # noqa: E731
        vol_b = torch.full((N_VOX[0] * N_VOX[1] * N_VOX[2],), -1, dtype=torch.int32)
        vol_b[lin(gl[valid])] = torch.nonzero(valid).squeeze(1).to(torch.int32)
        row_b = vol_b[lin(upd)]
        svol = torch.zeros(N_VOX, dtype=torch.int32)
        cb = coords[:, 1:]
        svol[cb[:, 0], cb[:, 1], cb[:, 2]] = info["panoptic_seg"][0]
        seg_u = svol[upd[:, 0], upd[:, 1], upd[:, 2]]
        ni, ns = fuse.panoptic_fusion(scale=2, global_valid=valid, relative_origin=rel.tolist(),
                                      panoptic_info={"panoptic_seg": [seg_u, info["panoptic_seg"][1]]},
                                      current_coords=upd, row_b=row_b)
        assert torch.equal(ni, state.I[-u:]) and torch.equal(ns, state.S[-u:]), frag
