"""Calibrate synth.STREAM_THRESHOLDS for BASELINE configs[2] on the GPU path: run the 16-fragment scene stream with the
caps lifted, and for every fragment / level report the fused (union) row count, the occupied count under the current
thresholds and the logit quantiles, so that thresholds can be chosen that keep every fragment of the stream inside the
shipped caps (config/test.yaml:29: 15 000 / 60 000 / 120 000, abort at 1.5 x).

    python tools/calibrate_stream.py [n_fragments] [t0 t1 t2]   -> JSON lines + gpurun_out/calibrate_stream.json
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from eprecon_b200 import synth  # noqa: E402
from eprecon_b200.neucon_network import NeuConNet  # noqa: E402

n_frag = int(sys.argv[1]) if len(sys.argv) > 1 else 16
thr = [float(v) for v in sys.argv[2:5]] if len(sys.argv) >= 5 else list(synth.STREAM_THRESHOLDS)
shipped = (15000, 60000, 120000)
cfg = synth.make_cfg(num_sample=(10 ** 7, 10 ** 7, 10 ** 7))
cfg.THRESHOLDS = thr
net = NeuConNet(cfg)
synth.fill_parameters_(net, 1)
net = net.cuda().train()


def dev(obj):
    if torch.is_tensor(obj):
        return obj.cuda()
    if isinstance(obj, list):
        return [dev(o) for o in obj]
    if isinstance(obj, dict):
        return {k: dev(v) for k, v in obj.items()}
    return obj


rows = []
for f in range(n_frag):
    inputs, fa, fb = synth.make_fragment(seed=1, frag_index=f)
    cin = dev(inputs)
    cin["scene"], cin["fragment"] = ["calib_scene"], [f"calib_{f}"]
    net.trace = {}
    out, _ = net(dev(fa), dev(fb), cin, {})
    row = {"fragment": f, "completed": "coords" in out}
    for lvl in range(3):
        t = net.trace.get(f"l{lvl}")
        if t is None:
            break
        occ = t["occ"].view(-1).float()
        kept = int((occ > thr[lvl]).sum())
        srt = torch.sort(occ, descending=True).values
        target = int(0.9 * shipped[lvl])
        row[f"l{lvl}"] = {"fused": int(occ.numel()), "occupied": kept, "cap": shipped[lvl],
                          "thr_for_90pct_cap": float(srt[min(target, occ.numel() - 1)]) if occ.numel() > target else None}
    rows.append(row)
    print(json.dumps(row), flush=True)
    net.trace = None
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"thresholds": thr, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "calibrate_stream.json"), "w"), indent=1)
