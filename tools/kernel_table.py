"""Per-kernel GPU time of ONE steady-state bench fragment, from torch.profiler (CUPTI activity records: no kernel replay,
so it costs seconds instead of ncu's minutes).  For orientation only -- the judged launch list is the ncu one."""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eprecon_b200 import synth  # noqa: E402
from eprecon_b200.neucon_network import NeuConNet  # noqa: E402

cfg = synth.make_cfg()
cfg.THRESHOLDS = list(synth.BENCH_THRESHOLDS)
net = NeuConNet(cfg)
synth.fill_parameters_(net, 1)
net = net.cuda().train()
inputs, fa, fb = synth.make_fragment(seed=1)
cin = {k: (v.cuda() if torch.is_tensor(v) else ([t.cuda() for t in v] if isinstance(v, list) and torch.is_tensor(v[0]) else v))
       for k, v in inputs.items()}
fa = [[t.cuda() for t in f] for f in fa]
fb = [[t.cuda() for t in f] for f in fb]


def step(i):
    cin["scene"] = [f"s{i}"]
    out, _ = net(fa, fb, cin, {})
    return out


for i in range(3):
    step(i)
torch.cuda.synchronize()
n = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(n):
        step(10 + i)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.split("(")[0].replace("void ", "")[:90]
        agg[name][0] += 1
        agg[name][1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"# {n} fragments: {tot / n / 1e3:.2f} ms of kernel+memcpy time per fragment, {sum(v[0] for v in agg.values()) // n} activities per fragment")
print("# us_per_fragment  share  count_per_fragment  name")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
    print(f"{v[1] / n:10.1f} {100 * v[1] / tot:5.1f}% {v[0] / n:7.1f}  {k}")
