"""Generate tests/golden/panoptic_fusion_small.npz by running the UNMODIFIED reference
GRUFusion(direct_substitute=True).forward (models/gru_fusion.py:259-394 with panoptic_fusion :133-193,
compute_overlap :116-131, update_map :195-215, save_mesh :217-257) on three overlapping synthetic fragments.

    python tests/golden/make_golden_panoptic_fusion.py      # needs /root/reference

Direct-substitute fusion is plain ATen (no sparse convolution), so the fixture pins oracle/restate.py's
`gru_fusion(..., panoptic_info=...)` directly against the reference.  Inputs are regenerated in the tests by
`fusion_inputs()` below (deterministic integer / torch.Generator arithmetic) and stored in the fixture as well.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from eprecon_b200 import synth  # noqa: E402

N_VOX = (48, 48, 48)
REGION_CLASS = {0: 1, 1: 5, 2: 7, 3: 5, 4: 2, 5: 9, 6: 1, 7: 5}     # 1, 2 = stuff (wall, floor); the rest are things


def fusion_inputs(frag):
    """(inputs, coords int64 [N,4], tsdf [N,1], panoptic_info) of fragment `frag` of one synthetic scene."""
    inputs, _, _ = synth.make_fragment(seed=1, image_hw=(240, 320), n_vox=N_VOX, frag_index=frag, with_features=False)
    g = torch.Generator().manual_seed(100 + frag)
    occ = inputs["occ_list"][0][0]
    c = torch.nonzero(occ)
    c = c[torch.randperm(len(c), generator=g)[: (2 * len(c)) // 3]]
    c = c[torch.argsort((c[:, 0] * 64 + c[:, 1]) * 64 + c[:, 2])]
    coords = torch.cat([torch.zeros(len(c), 1, dtype=torch.long), c], 1)
    tsdf = torch.rand(len(c), 1, generator=g) * 2.4 - 1.2               # some |tsdf| >= 1 rows (inactive)
    rel = ((inputs["vol_origin_partial"][0] - inputs["vol_origin"][0]) / 0.04).long()
    gc = c + rel
    region = (torch.div(gc[:, 0], 20, rounding_mode="floor") % 2) + 2 * (torch.div(gc[:, 1], 20, rounding_mode="floor") % 2) \
        + 4 * (torch.div(gc[:, 2], 30, rounding_mode="floor") % 2)
    drop = torch.rand(len(c), generator=g) < 0.1                        # unlabelled voxels (segment id 0)
    seg = torch.zeros(len(c), dtype=torch.int32)
    info = []
    for r in sorted(set(region.tolist())):
        cls = REGION_CLASS[int(r)]
        m = (region == r) & ~drop
        if int(m.sum()) == 0:
            continue
        prev = [d for d in info if d["category_id"] == cls and not d["isthing"]]
        if cls in (1, 2) and prev:                                      # stuff classes share one segment id
            seg[m] = prev[0]["id"]
            continue
        info.append({"id": len(info) + 1, "isthing": cls not in (1, 2), "category_id": cls})
        seg[m] = len(info)
    return inputs, coords, tsdf, {"panoptic_seg": [seg, info]}


def canonical(C, *cols):
    key = (C[:, 0] * 100000 + C[:, 1]) * 100000 + C[:, 2]
    o = torch.argsort(key)
    return [C[o]] + [x.reshape(len(C), -1)[o] for x in cols]


def main():
    from oracle import ref_import
    ns = ref_import.load()
    cfg = synth.make_cfg(n_vox=N_VOX)
    fuse = ns.gru.GRUFusion(cfg, direct_substitute=True, trianing=False)
    g, outputs = {}, {}
    for frag in range(3):
        inputs, coords, tsdf, info = fusion_inputs(frag)
        g[f"f{frag}_coords"] = coords.numpy().astype(np.int16)
        g[f"f{frag}_tsdf"] = tsdf.numpy()
        g[f"f{frag}_seg"] = info["panoptic_seg"][0].numpy().copy()
        g[f"f{frag}_info"] = np.asarray([[d["id"], int(d["isthing"]), d["category_id"]] for d in info["panoptic_seg"][1]], dtype=np.int32)
        ref_in = {k: v for k, v in inputs.items() if k not in ("occ_list", "tsdf_list")}
        with torch.no_grad():
            outputs = fuse(coords, tsdf, ref_in, 2, outputs, save_mesh=True, panoptic_infos=[info])
        C, F, I, S = canonical(fuse.global_volume[2].C.long(), fuse.global_volume[2].F, fuse.global_instance, fuse.global_semantic)
        g[f"f{frag}_gC"], g[f"f{frag}_gF"] = C.numpy().astype(np.int32), F.numpy().astype(np.float32)
        g[f"f{frag}_gI"], g[f"f{frag}_gS"] = I.numpy().astype(np.int32), S.numpy().astype(np.int32)
        g[f"f{frag}_scene_shape"] = np.asarray(outputs["scene_tsdf"][-1].shape)
        g[f"f{frag}_scene_instance_hist"] = np.bincount(outputs["scene_instance"][-1].long().flatten().numpy())
        g[f"f{frag}_scene_semantic_hist"] = np.bincount(outputs["scene_semantic"][-1].long().flatten().numpy())
        print(frag, "global rows", len(C), "instances", np.unique(g[f"f{frag}_gI"]).tolist(), "semantics", np.unique(g[f"f{frag}_gS"]).tolist())
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "panoptic_fusion_small.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
