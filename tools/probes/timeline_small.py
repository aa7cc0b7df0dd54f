"""Timeline + timing of SMALL sparse convs through the fused entry point (split-K tickets + fused BN finalisation)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from eprecon_b200 import _lib, ops  # noqa: E402
from eprecon_b200.sparse import VoxelSet  # noqa: E402
from probe_hl import slab_set, timeit  # noqa: E402

L = _lib.lib()
for n_side, cin, cout in ((5, 128, 128), (11, 96, 96), (11, 64, 64), (26, 96, 96)):
    coords = slab_set(n_side).cuda()
    nbr = VoxelSet(coords, 1).kmap_k3()
    m = coords.shape[0]
    x = torch.zeros(m, ops.ceil4(cin)); x[:, :cin] = torch.randn(m, cin)
    W = torch.zeros(27, cin, ops.ceil4(cout)); W[:, :, :cout] = torch.randn(27, cin, cout) / (27 * cin) ** 0.5
    xc, Wc = x.cuda(), W.cuda()
    w_hl, npad = ops._hl_weights(Wc, cout)
    x_hl = ops.hl_split(xc, cin)
    o = torch.empty((m, ops.ceil4(cout)), dtype=torch.float32, device="cuda")
    part = torch.empty((L.ep_spconv_num_row_tiles(m), 2, cout), dtype=torch.float32, device="cuda")
    ss = torch.empty((2, cout), dtype=torch.float32, device="cuda")
    gamma = torch.ones(cout, device="cuda"); beta = torch.zeros(cout, device="cuda")
    wsb = L.ep_spconv_hl_workspace_bytes(m, npad, 27)
    ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device="cuda")
    ctr = torch.zeros(1024, dtype=torch.int32, device="cuda")
    st = ops.stream_ptr()

    def fused():
        _lib.check(L.ep_spconv_hl_fused_fwd(x_hl.data_ptr(), m, cin, nbr.data_ptr(), 27, w_hl.data_ptr(), npad, cout, 0, o.data_ptr(),
                                            o.stride(0), m, part.data_ptr(), ws.data_ptr(), wsb, 0, ctr.data_ptr(), 1024,
                                            gamma.data_ptr(), beta.data_ptr(), 1e-5, ss.data_ptr(), st), "fused")

    def unfused():
        _lib.check(L.ep_spconv_hl_fwd(x_hl.data_ptr(), m, cin, nbr.data_ptr(), 27, w_hl.data_ptr(), npad, cout, 0, o.data_ptr(),
                                      o.stride(0), m, part.data_ptr(), ws.data_ptr(), wsb, 0, st), "unfused")
    def fused_nobn():
        _lib.check(L.ep_spconv_hl_fused_fwd(x_hl.data_ptr(), m, cin, nbr.data_ptr(), 27, w_hl.data_ptr(), npad, cout, 0, o.data_ptr(),
                                            o.stride(0), m, part.data_ptr(), ws.data_ptr(), wsb, 0, ctr.data_ptr(), 1024,
                                            0, 0, 0.0, 0, st), "fused_nobn")

    def fused_nopart():
        _lib.check(L.ep_spconv_hl_fused_fwd(x_hl.data_ptr(), m, cin, nbr.data_ptr(), 27, w_hl.data_ptr(), npad, cout, 0, o.data_ptr(),
                                            o.stride(0), m, 0, ws.data_ptr(), wsb, 0, ctr.data_ptr(), 1024,
                                            0, 0, 0.0, 0, st), "fused_nopart")
    print("fused without BN finalisation: %.1f us; without BN partials at all: %.1f us" % (timeit(fused_nobn, 50), timeit(fused_nopart, 50)))
    t_f, t_u = timeit(fused, 50), timeit(unfused, 50)
    # host cost of one call (no GPU wait)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200):
        fused()
    host = (time.perf_counter() - t0) / 200 * 1e6
    torch.cuda.synchronize()
    dbg = torch.zeros(400, dtype=torch.int64, device="cuda")
    L.ep_hl_set_timeline(dbg.data_ptr()); fused(); torch.cuda.synchronize(); L.ep_hl_set_timeline(0)
    d = dbg.cpu().tolist(); t0_ = d[386]
    rows = [[v - t0_ if v else None for v in d[6 * t:6 * t + 6]] for t in range(64) if d[6 * t + 2]]
    print("final-CTA stamps (clk since the sampled CTA's entry): z", d[389], "[ns since the sampled CTA's entry] sampled CTA done", d[395] - d[394], "ticket won", d[390] - d[394], "reduce done", d[391] - d[394], "bn ticket won", d[392] - d[394], "finalise done", d[393] - d[394])
    print(f"m {m} cin {cin} cout {cout}: fused {t_f:.1f} us, unfused(+reduce kernel) {t_u:.1f} us, host issue {host:.1f} us/call; "
          f"CTA timeline: first stage {rows[0] if rows else None} last {rows[-1] if rows else None} mainloop_done {d[384] - t0_} epilogue_done {d[385] - t0_}")
