"""Generate tests/golden/*.npz by running the UNMODIFIED reference (over oracle/shims) in the build container.

    python tests/golden/make_golden.py          # needs /root/reference; writes tests/golden/neucon_small.npz

The fixture pins oracle/restate.py (the travelling CPU restatement) against the reference's own modules:
NeuConNet.forward on one small synthetic fragment (N_VOX 64^3, 320x240 images, 9 views, seed 1, weights from
eprecon_b200.synth.fill_parameters_).  The panoptic decoder (out of scope) is replaced by a stub so
the run takes ~1 min; everything up to outputs['coords'] / outputs['tsdf'] is the reference's code.
Per stage we store shapes, float64 sums and a strided row sample; the final sparse TSDF is stored in full.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from eprecon_b200 import synth  # noqa: E402

SMALL = dict(n_vox=(64, 64, 64), image_hw=(240, 320), n_views=9, seed=1)
OCC_FRACTION = [0.9, 0.8, 0.6]


def sample_rows(x, k=256):
    x = x.detach().cpu()
    if x.dtype == torch.bool:
        x = x.to(torch.uint8)
    idx = torch.linspace(0, x.shape[0] - 1, min(k, x.shape[0])).long()
    return idx.numpy(), x[idx].numpy()


def main():
    ns = ref_import.load()
    cfg = synth.make_cfg(n_vox=SMALL["n_vox"])
    torch.manual_seed(1)
    np.random.seed(1)
    net = ns.neucon.NeuConNet(cfg)
    synth.fill_parameters_(net, 1)
    net.train()  # the reference evaluates in train mode (main.py:357)
    # stub the out-of-scope panoptic decoder + post-processing
    net.panoptic.forward = lambda panoptic_features, panoptic_coords, mask_features, spitial_shape: {
        "pred_logits": torch.zeros(1, 80, 21), "pred_masks": torch.zeros(1, 80, mask_features.shape[-1]), "aux_outputs": []}
    ns.neucon.panoptic_post = lambda out: {"panoptic_seg": [torch.zeros(out["pred_masks"].shape[-1]), []]}

    rec = {}

    def hook(name):
        def f(mod, inp, out):
            rec[name] = out
        return f

    def occ_hook(i):
        def f(mod, inp, out):
            q = torch.quantile(out.flatten(), 1 - OCC_FRACTION[i]).item()
            cfg.THRESHOLDS[i] = round(q, 3)   # calibrated on the reference, stored in the fixture
            rec[f"occ{i}"] = out
        return f

    net.initialization.register_forward_hook(hook("init"))
    net.gru_fusion.register_forward_hook(lambda m, i, o: rec.setdefault("gru", []).append(o))
    for i in range(3):
        net.back_projection[i].register_forward_hook(hook(f"bp{i}"))
        net.sp_convs[i].register_forward_hook(hook(f"spv{i}"))
        net.tsdf_preds[i].register_forward_hook(hook(f"tsdf{i}"))
        net.occ_preds[i].register_forward_hook(occ_hook(i))
    for p in range(3):
        net.panoptic_preds[p].register_forward_hook(hook(f"pano{p}"))
    net.panoptic_feat_fusion.mask_feat_extraction_2.register_forward_hook(hook("mask_feats"))
    inputs, fa, fb = synth.make_fragment(seed=SMALL["seed"], n_views=SMALL["n_views"], image_hw=SMALL["image_hw"],
                                         n_vox=SMALL["n_vox"])
    with torch.no_grad():
        out, _ = net(fa, fb, inputs, {})
    assert "coords" in out, "reference forward did not complete"
    g = {"thresholds": np.asarray(cfg.THRESHOLDS, dtype=np.float64), "n_vox": np.asarray(SMALL["n_vox"]),
         "image_hw": np.asarray(SMALL["image_hw"]), "n_views": np.asarray(SMALL["n_views"]), "seed": np.asarray(SMALL["seed"]),
         "final_coords": out["coords"].numpy().astype(np.int16), "final_tsdf": out["tsdf"].numpy().astype(np.float32)}

    def put(name, x):
        x = x.detach()
        g[name + "_shape"] = np.asarray(x.shape)
        g[name + "_sum"] = np.asarray(x.double().sum().item())
        g[name + "_abssum"] = np.asarray(x.double().abs().sum().item())
        g[name + "_idx"], g[name + "_rows"] = sample_rows(x)

    put("init_occ", rec["init"][0])
    put("init_coords", rec["init"][1])
    g["init_count_hist"] = np.bincount(rec["init"][2].long().numpy(), minlength=10)
    for i in range(3):
        put(f"bp{i}_feat", rec[f"bp{i}"][0])
        put(f"bp{i}_coords", rec[f"bp{i}"][1])
        g[f"bp{i}_count_hist"] = np.bincount(rec[f"bp{i}"][4].long().numpy(), minlength=10)
        put(f"spv{i}", rec[f"spv{i}"])
        put(f"gru{i}_coords", rec["gru"][i][0])
        put(f"gru{i}_values", rec["gru"][i][1])
        put(f"tsdf{i}", rec[f"tsdf{i}"])
        put(f"occ{i}", rec[f"occ{i}"])
    for p in range(3):
        put(f"pano{p}", rec[f"pano{p}"])
    put("mask_feats", rec["mask_feats"])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "neucon_small.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path), "bytes; thresholds", cfg.THRESHOLDS,
          "sizes", [int(rec[f"occ{i}"].shape[0]) for i in range(3)], "final", out["coords"].shape)


if __name__ == "__main__":
    main()
