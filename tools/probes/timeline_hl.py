"""clock64 timeline of one CTA of spconv_hl_cp_kernel (level-2 sizes): where do the producer and the MMA thread wait?"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from eprecon_b200 import _lib, ops  # noqa: E402
from eprecon_b200.sparse import VoxelSet  # noqa: E402
from probe_hl import slab_set  # noqa: E402

L = _lib.lib()
coords = slab_set(182).cuda()
nbr = VoxelSet(coords, 1).kmap_k3()
m = coords.shape[0]
res = {}
for cin, cout in ((24, 24), (74, 8), (96, 96)):
    x = torch.zeros(m, ops.ceil4(cin)); x[:, :cin] = torch.randn(m, cin)
    W = torch.zeros(27, cin, ops.ceil4(cout)); W[:, :, :cout] = torch.randn(27, cin, cout) / (27 * cin) ** 0.5
    xc, Wc = x.cuda(), W.cuda()
    w_hl, npad = ops._hl_weights(Wc, cout)
    x_hl = ops.hl_split(xc, cin)
    o = torch.empty((m, ops.ceil4(cout)), dtype=torch.float32, device="cuda")
    dbg = torch.zeros(400, dtype=torch.int64, device="cuda")
    for it in range(3):
        if it == 2:
            L.ep_hl_set_timeline(dbg.data_ptr())
        _lib.check(L.ep_spconv_hl_fwd(x_hl.data_ptr(), m, cin, nbr.data_ptr(), 27, w_hl.data_ptr(), npad, cout, 0, o.data_ptr(),
                                      o.stride(0), m, 0, 0, 0, 0, ops.stream_ptr()), "hl")
    torch.cuda.synchronize()
    L.ep_hl_set_timeline(0)
    d = dbg.cpu().tolist()
    t0 = d[386]
    rows = [[v - t0 if v else None for v in d[6 * t:6 * t + 6]] for t in range(64) if d[6 * t + 2]]
    res[f"{cin}->{cout}"] = {"stages": rows, "mainloop_done": d[384] - t0, "epilogue_done": d[385] - t0}
    print(cin, cout, "stages", len(rows), "mainloop_done", d[384] - t0, "epilogue_done", d[385] - t0)
    for t, r in enumerate(rows[:40]):
        print(t, r)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "timeline_hl.json"), "w"))
