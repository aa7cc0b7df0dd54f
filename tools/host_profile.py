"""cProfile of the host side of NeuConNet.forward (where do the ~800 launches/step spend Python time?)."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eprecon_b200 import synth  # noqa: E402
from eprecon_b200.neucon_network import NeuConNet  # noqa: E402

cfg = synth.make_cfg()
cfg.THRESHOLDS = list(synth.BENCH_THRESHOLDS)
net = NeuConNet(cfg)
synth.fill_parameters_(net, 1)
net = net.cuda()
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
t0 = time.perf_counter()
for i in range(5):
    step(10 + i)
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile()
pr.enable()
for i in range(5):
    step(20 + i)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(45)
st.sort_stats("cumulative").print_stats(60)
