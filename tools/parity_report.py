"""Per-stage parity report of the CUDA path against the CPU oracle on the small golden configuration, for every
sparse-conv implementation (fp32 FFMA, tcgen05 3xTF32, tcgen05 TF32).  Prints one JSON line per implementation.
`--full`: the same report on the FULL-size bench fragment (BASELINE configs[1]: 96^3, 9 x 640 x 480, ~210 k level-2
candidates) -- the oracle takes ~12-22 s there, so this is the end-to-end parity check at the benchmarked size."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eprecon_b200 import ops, synth  # noqa: E402
from eprecon_b200.neucon_network import NeuConNet  # noqa: E402
from oracle import restate  # noqa: E402


def rel(a, b):
    a, b = a.detach().cpu().float(), b.detach().cpu().float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def elem(a, b, floor):
    """worst |a-b| / (1e-3 * max(|b|, floor * max|b|)): <= 1 means every entry at least `floor` of the tensor's scale is
    within 1e-3 of ITSELF (the smaller ones are bounded by 1e-3 * floor * scale)."""
    a, b = a.detach().cpu().float(), b.detach().cpu().float()
    bound = 1e-3 * torch.maximum(b.abs(), floor * b.abs().max())
    return float(((a - b).abs() / bound).max())


def main():
    full = "--full" in sys.argv
    if full:
        # BASELINE configs[1] at its FULL size (96^3, 9 x 640 x 480, the bench fragment): the oracle needs ~12-22 s of CPU
        cfg = synth.make_cfg()
        cfg.THRESHOLDS = list(synth.BENCH_THRESHOLDS)
        frag = dict(seed=1)
    else:
        g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "neucon_small.npz"))
        n_vox = tuple(int(v) for v in g["n_vox"])
        cfg = synth.make_cfg(n_vox=n_vox)
        cfg.THRESHOLDS = [float(v) for v in g["thresholds"]]
        frag = dict(seed=1, image_hw=(240, 320), n_vox=n_vox)
    net = NeuConNet(cfg)
    sd = synth.synthetic_state_dict(net, 1)
    inputs, fa, fb = synth.make_fragment(**frag)
    ot = {}
    with torch.no_grad():
        oout = restate.neucon_forward(sd, cfg, fa, fb, inputs, restate.FusionState(), trace=ot)
    net = net.cuda()
    cin = {k: (v.cuda() if torch.is_tensor(v) else ([t.cuda() for t in v] if isinstance(v, list) and torch.is_tensor(v[0]) else v))
           for k, v in inputs.items()}
    fa_c, fb_c = [[t.cuda() for t in f] for f in fa], [[t.cuda() for t in f] for f in fb]
    impls = [a for a in sys.argv[1:] if not a.startswith("--")] or ["ffma", "tf32x3", "hl", "tf32"]
    for impl in impls:
        ops.SPCONV_IMPL = impl
        cin["scene"] = [f"scene_{impl}"]
        net.trace, net.teacher = {}, ot
        out, _ = net(fa_c, fb_c, cin, {})
        t = net.trace
        rep = {"impl": impl, "config": "full 96^3" if full else "golden 64^3", "init_occ": rel(t["init"]["occ"], ot["init"]["occ"])}
        for lv in range(3):
            a, b = t[f"l{lv}_pre_gru"], ot[f"l{lv}_pre_gru"]
            rep[f"l{lv}_coords_exact"] = bool(torch.equal(a["coords"].cpu(), b["coords"].int()))
            rep[f"l{lv}_spvcnn"] = rel(a["spvcnn"], b["spvcnn"])
            a, b = t[f"l{lv}"], ot[f"l{lv}"]
            rep[f"l{lv}_union_exact"] = bool(torch.equal(a["coords"].cpu().long(), b["coords"]))
            rep[f"l{lv}_gru"] = rel(a["feat_all"], b["feat_all"])
            rep[f"l{lv}_tsdf"] = rel(a["tsdf"], b["tsdf"])
            rep[f"l{lv}_occ"] = rel(a["occ"], b["occ"])
            rep[f"l{lv}_mask_flips"] = int((a["occupancy"].cpu() != b["occupancy"]).sum())
        rep["final_coords_exact"] = bool(torch.equal(out["coords"].cpu(), oout["coords"]))
        rep["final_tsdf"] = rel(out["tsdf"], oout["tsdf"])
        # elementwise view (worst ratio to the bound) at three floors, over every float stage
        pairs = [("init_occ", t["init"]["occ"], ot["init"]["occ"]), ("final_tsdf", out["tsdf"], oout["tsdf"])]
        for lv in range(3):
            pairs += [(f"l{lv}_spvcnn", t[f"l{lv}_pre_gru"]["spvcnn"], ot[f"l{lv}_pre_gru"]["spvcnn"]),
                      (f"l{lv}_gru", t[f"l{lv}"]["feat_all"], ot[f"l{lv}"]["feat_all"]),
                      (f"l{lv}_tsdf", t[f"l{lv}"]["tsdf"], ot[f"l{lv}"]["tsdf"]), (f"l{lv}_occ", t[f"l{lv}"]["occ"], ot[f"l{lv}"]["occ"])]
        for floor in (0.01, 0.05, 0.25):
            worst = max((elem(a, b, floor), name) for name, a, b in pairs)
            rep[f"elementwise_worst@floor{floor}"] = [round(worst[0], 3), worst[1]]
        print(json.dumps(rep))


if __name__ == "__main__":
    main()
