"""Where the panoptic branch's time goes: torch.profiler (CUPTI activity records, no replay) over ONE fragment with
`with_panoptic`, split into the TSDF path and the panoptic tail by NVTX-free bookkeeping: kernels launched after
outputs['coords'] exists are attributed to the panoptic branch.  Orientation only (the judged lists are the ncu ones)."""
import collections
import os
import sys
import time

import torch
from torch.profiler import ProfilerActivity, profile, record_function

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eprecon_b200 import synth  # noqa: E402
from eprecon_b200.neucon_network import NeuConNet  # noqa: E402

cfg = synth.make_cfg()
cfg.THRESHOLDS = list(synth.BENCH_THRESHOLDS)
net = NeuConNet(cfg)
synth.fill_parameters_(net, 1)
net = net.cuda().train()
net.with_panoptic = True
inputs, fa, fb = synth.make_fragment(seed=1)
cin = {k: (v.cuda() if torch.is_tensor(v) else ([t.cuda() for t in v] if isinstance(v, list) and torch.is_tensor(v[0]) else v))
       for k, v in inputs.items()}
fa = [[t.cuda() for t in f] for f in fa]
fb = [[t.cuda() for t in f] for f in fb]

# wrap the two stages of the panoptic tail so that their kernels can be told apart
orig_prepare, orig_decode = net.panoptic_prepare, net.panoptic_decode


def prepare(*a, **k):
    with record_function("PANO_PREPARE"):
        return orig_prepare(*a, **k)


def decode(*a, **k):
    with record_function("PANO_DECODE"):
        return orig_decode(*a, **k)


net.panoptic_prepare, net.panoptic_decode = prepare, decode


def step(i):
    cin["scene"] = [f"s{i}"]
    return net(fa, fb, cin, {})[0]


for i in range(3):
    step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(3):
    step(10 + i)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 3
net.with_panoptic = False
for i in range(2):
    step(20 + i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(3):
    step(30 + i)
torch.cuda.synchronize()
wall_tsdf = (time.perf_counter() - t0) / 3
net.with_panoptic = True
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    step(40)
    torch.cuda.synchronize()
print(f"# wall per fragment: {wall * 1e3:.2f} ms with the panoptic head, {wall_tsdf * 1e3:.2f} ms without")
ev = prof.events()
spans = {}
for e in ev:
    if e.name in ("PANO_PREPARE", "PANO_DECODE") and e.device_type == torch.autograd.DeviceType.CPU:
        spans[e.name] = (e.time_range.start, e.time_range.end)
agg = {k: collections.defaultdict(lambda: [0, 0.0]) for k in ("PANO_PREPARE", "PANO_DECODE")}
# attribute CUDA kernels to a stage through the CPU launch op that encloses them (correlation by launch time)
for e in ev:
    if e.device_type != torch.autograd.DeviceType.CPU or not e.kernels:
        continue
    for stage, (a, b) in spans.items():
        if a <= e.time_range.start <= b:
            for k in e.kernels:
                n = k.name.split("(")[0].replace("void ", "")[:80]
                agg[stage][n][0] += 1
                agg[stage][n][1] += k.duration
for stage in agg:
    tot = sum(v[1] for v in agg[stage].values())
    print(f"# {stage}: {tot / 1e3:.2f} ms of kernel time, {sum(v[0] for v in agg[stage].values())} launches, "
          f"host span {(spans[stage][1] - spans[stage][0]) / 1e3:.2f} ms" if stage in spans else f"# {stage}: not seen")
    for k, v in sorted(agg[stage].items(), key=lambda kv: -kv[1][1])[:18]:
        print(f"{v[1]:10.1f} us {v[0]:5d}  {k}")
