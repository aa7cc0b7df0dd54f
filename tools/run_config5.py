"""BASELINE configs[4] ("high-res stress"): 18-view window, 960x720 frames, 128^3 finest grid, TSDF path.  No oracle at
this size (minutes of CPU); checks that the path completes, reports sizes and fragments/s on one GPU."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eprecon_b200 import synth  # noqa: E402
from eprecon_b200.neucon_network import NeuConNet  # noqa: E402

n_vox = (128, 128, 128)
cfg = synth.make_cfg(n_vox=n_vox, num_sample=(40000, 160000, 400000))
cfg.THRESHOLDS = list(synth.BENCH_THRESHOLDS)
net = NeuConNet(cfg)
synth.fill_parameters_(net, 1)
net = net.cuda()
inputs, fa, fb = synth.make_fragment(seed=1, n_views=18, image_hw=(720, 960), n_vox=n_vox)
cin = {k: (v.cuda() if torch.is_tensor(v) else ([t.cuda() for t in v] if isinstance(v, list) and torch.is_tensor(v[0]) else v))
       for k, v in inputs.items()}
fa = [[t.cuda() for t in f] for f in fa]
fb = [[t.cuda() for t in f] for f in fb]


def step(i):
    cin["scene"] = [f"s{i}"]
    return net(fa, fb, cin, {})[0]


out = step(0)
ok = "coords" in out
for i in range(2):
    step(1 + i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 5
for i in range(K):
    out = step(10 + i)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(json.dumps({"config": "18 views, 960x720, 128^3", "completed": ok, "sizes": net.last_sizes, "ms_per_fragment": ms,
                  "fragments_per_s": 1e3 / ms, "tsdf_finite": bool(torch.isfinite(out["tsdf"]).all()) if ok else None,
                  "final_voxels": int(out["coords"].shape[0]) if ok else 0,
                  "max_mem_GB": torch.cuda.max_memory_allocated() / 2 ** 30}))
