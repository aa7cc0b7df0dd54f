"""Dense per-row linears of the bench fragment (K == 1): the shipped kernel (EPRECON_LINEAR: 2 = linear_mma_kernel, default; 1 =
linear_rows_kernel) vs the gather-GEMM tile kernel (identity table).

    python tools/probes/probe_linear.py > gpurun_out/r02_probe_linear_rows.json
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from eprecon_b200 import _lib, ops  # noqa: E402

L = _lib.lib()
SHAPES = [(209800, 24, 96), (209800, 96, 24), (209800, 32, 24), (209800, 8, 32), (32776, 192, 48), (209800, 48, 24), (37137, 32, 32),
          (105000, 48, 192), (105000, 192, 48)]


FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, flush, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            FLUSH.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


rows = []
for m, cin, cout in SHAPES:
    g = torch.Generator().manual_seed(cin)
    x = torch.randn(m, ops.ceil4(cin), generator=g).cuda()
    W = (torch.randn(1, cin, ops.ceil4(cout), generator=g) / cin ** 0.5).cuda()
    out = torch.empty(m, ops.ceil4(cout), device="cuda")
    part = torch.empty(L.ep_spconv_num_row_tiles(m), 2, cout, device="cuda")
    nbr = torch.arange(m, dtype=torch.int32, device="cuda").view(m, 1).contiguous()
    st = ops.stream_ptr()

    def run(nb):
        _lib.check(L.ep_spconv_fwd(x.data_ptr(), x.stride(0), cin, nb, 1, W.data_ptr(), W.shape[2], cout, 0, out.data_ptr(),
                                   out.stride(0), m, part.data_ptr(), st), "fwd")
    b = (m * cin + m * cout) * 4
    rec = {"m": m, "cin": cin, "cout": cout, "MB": round(b / 1e6, 1)}
    for tag, flush in (("warm", False), ("flushed", True)):      # warm: operands L2-resident, as right after their producer
        new = timed(lambda: run(0), flush)
        old = timed(lambda: run(nbr.data_ptr()), flush)
        rec.update({f"linear_us_{tag}": round(new, 1), f"tile_us_{tag}": round(old, 1), f"linear_GBs_{tag}": round(b / new / 1e3)})
    rows.append(rec)
print(json.dumps({"linear": rows}, indent=1))
