"""Generate tests/golden/mask3dformer_small.npz by running the UNMODIFIED reference panoptic decoder
(models/mask3dformer.py: MultiScaleMaskedTransformerDecoder.forward + panoptic_post / panoptic_inference) on CPU.

    python tests/golden/make_golden_mask3dformer.py      # needs /root/reference

The decoder is plain ATen (no torchsparse / spconv underneath), so this fixture pins oracle/restate.py's
`mask3dformer`, `nearest_fine_index`, `fourier_positions` and `panoptic_inference` directly against the reference.
Inputs: a synthetic 2-manifold shell in a 24^3 volume (level-2 voxels in random order, level-1 / level-0 voxels = the
parents / grand-parents, as the reference's level alignment leaves them), N(0,1) features, weights from
eprecon_b200.synth.fill_parameters_(prefix="panoptic.").
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from eprecon_b200 import synth  # noqa: E402

DIM = 24


def shell_inputs(seed=3, dim=DIM):
    g = torch.Generator().manual_seed(seed)
    ax = torch.arange(dim, dtype=torch.float32)
    X, Y, Z = torch.meshgrid(ax, ax, ax, indexing="ij")
    sdf = torch.minimum(Z - 2.5, torch.minimum(dim - 3.5 - X, ((X - 8) ** 2 + (Y - 12) ** 2 + (Z - 6) ** 2).sqrt() - 5.0))
    fine = torch.nonzero(sdf.abs() < 1.0)
    fine = fine[torch.randperm(len(fine), generator=g)]
    mid = torch.unique(torch.floor_divide(fine, 2) * 2, dim=0)
    mid = mid[torch.randperm(len(mid), generator=g)]
    top = torch.unique(torch.floor_divide(fine, 4) * 4, dim=0)
    top = top[torch.randperm(len(top), generator=g)]
    coords = [top, mid, fine]
    feats = [torch.randn(len(c), 48, generator=g) for c in coords]
    mask_features = torch.randn(len(fine), 48, generator=g)
    return coords, feats, mask_features


def main():
    ref_import.install()
    import importlib
    m3d = importlib.import_module("models.mask3dformer")
    torch.manual_seed(1)
    dec = m3d.MultiScaleMaskedTransformerDecoder(mask_classification=True, num_classes=20, hidden_dim=48, num_queries=80,
                                                 nheads=8, dim_feedforward=192, dec_layers=6, pre_norm=False, mask_dim=48)
    synth.fill_parameters_(dec, 1, prefix="panoptic.")
    dec.train()
    coords, feats, mask_features = shell_inputs()
    with torch.no_grad():
        out = dec(panoptic_features=[f.unsqueeze(0).permute(0, 2, 1) for f in feats],
                  panoptic_coords=[c.unsqueeze(0) for c in coords],
                  mask_features=mask_features.unsqueeze(0).permute(0, 2, 1), spitial_shape=(DIM, DIM, DIM))
        # the reference's own nearest-voxel indices (mask3dformer.py:361-368), recomputed with its two lines
        idx0 = torch.argmin(torch.cdist(coords[2].float(), coords[0].unsqueeze(0).float(), p=2), dim=1).view(-1)
        idx1 = torch.argmin(torch.cdist(coords[2].float(), coords[1].unsqueeze(0).float(), p=2), dim=1).view(-1)
        post = m3d.panoptic_post(out)
    g = {"dim": np.asarray(DIM), "n": np.asarray([len(c) for c in coords]),
         "coords0": coords[0].numpy().astype(np.int16), "coords1": coords[1].numpy().astype(np.int16),
         "coords2": coords[2].numpy().astype(np.int16),
         "feats0": feats[0].numpy(), "feats1": feats[1].numpy(), "feats2": feats[2].numpy(),
         "mask_features": mask_features.numpy(),
         "pred_logits": out["pred_logits"][0].numpy(), "pred_masks": out["pred_masks"][0].numpy(),
         "aux_logits": np.stack([a["pred_logits"][0].numpy() for a in out["aux_outputs"]]),
         "aux_masks_abssum": np.asarray([float(a["pred_masks"].double().abs().sum()) for a in out["aux_outputs"]]),
         "index0": idx0.numpy().astype(np.int32), "index1": idx1.numpy().astype(np.int32),
         "post_seg": post["panoptic_seg"][0].numpy()}
    # panoptic_inference on confident synthetic predictions (random-init logits rarely pass the 0.3 score threshold)
    gi = torch.Generator().manual_seed(11)
    for case in range(4):
        cls = torch.randn(80, 21, generator=gi) * 6.0
        msk = torch.randn(80, 1500, generator=gi) * 3.0 + torch.randn(80, 1, generator=gi) * 2.0
        if case == 2:
            cls[:, 0] += 20.0       # almost everything "no object": a handful of queries survive
        if case == 3:               # six confident queries on disjoint regions, repeated stuff classes (1, 2) -> merged ids
            cls[:, 0] += 40.0
            msk = torch.randn(80, 1500, generator=gi) - 6.0
            for k, lab in enumerate([1, 2, 1, 5, 2, 7]):
                cls[k, lab] += 80.0
                msk[k, 200 * k:200 * (k + 1)] += 12.0
        seg, info = m3d.panoptic_inference(cls, msk)
        g[f"pi{case}_cls"], g[f"pi{case}_msk"] = cls.numpy(), msk.numpy().astype(np.float32)
        g[f"pi{case}_seg"] = seg.numpy()
        g[f"pi{case}_info"] = np.asarray([[d["id"], int(d["isthing"]), d["category_id"]] for d in info], dtype=np.int32).reshape(-1, 3)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mask3dformer_small.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path), "bytes; sizes", [len(c) for c in coords],
          "segments", [len(g[f"pi{c}_info"]) for c in range(4)], "post nonzero", int((g["post_seg"] != 0).sum()))


if __name__ == "__main__":
    main()
