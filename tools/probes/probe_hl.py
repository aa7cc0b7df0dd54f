"""GPU probe for csrc/spconv_hl.cu: (1) TMA tile::gather4 semantics (row order in shared memory, SWIZZLE_128B chunk
permutation, out-of-bounds / negative rows -> zeros), (2) the conv against fp64 on random problems, (3) timing against the
round-1 tcgen05 kernel (3xTF32, register producers) on slab-shaped voxel sets in Morton order at the bench fragment's sizes.

    python tools/probes/probe_hl.py [--skip-timing]      -> prints + gpurun_out/probe_hl.json
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from eprecon_b200 import _lib, ops  # noqa: E402

OUT = {}


def probe_gather4():
    L = _lib.lib()
    m_in, nslab = 1000, 2
    x = torch.arange(m_in * nslab * 64, dtype=torch.int32).view(m_in, nslab * 64)
    vals = ((x // (nslab * 64)) * 7 + (x % (nslab * 64))).to(torch.int16).cuda().contiguous()   # unique-ish per (row, col)
    g = torch.Generator().manual_seed(0)
    rows = torch.randint(0, m_in, (128,), generator=g, dtype=torch.int32)
    rows[3] = -1
    rows[10] = m_in
    rows[11] = m_in + 77
    rows[64] = 2 ** 31 - 1
    rows[65] = -5
    rows[127] = rows[0]
    res = {}
    for slab in (0, 1):
        out = torch.zeros(16384, dtype=torch.uint8, device="cuda")
        status = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.check(L.ep_hl_probe_gather4(vals.data_ptr(), m_in, nslab, rows.cuda().data_ptr(), slab, out.data_ptr(),
                                         status.data_ptr(), ops.stream_ptr()), "probe")
        torch.cuda.synchronize()
        st = int(status.item())
        img = out.cpu().view(128, 8, 16)                   # [smem row][16-byte chunk][bytes]
        src = vals.cpu().view(torch.uint8).view(m_in, nslab, 8, 16)
        ok_rows, bad = 0, []
        for i in range(128):
            r = int(rows[i])
            want = src[r, slab] if 0 <= r < m_in else torch.zeros(8, 16, dtype=torch.uint8)
            perm = [c ^ (i & 7) for c in range(8)]       # SWIZZLE_128B: chunk c of row i lives at chunk c ^ (i % 8)
            got = img[i][perm]
            if torch.equal(got, want):
                ok_rows += 1
            else:
                bad.append((i, r, bool(torch.equal(img[i], want)), int((img[i] == 0).all())))
        res[f"slab{slab}"] = {"status": st, "rows_ok": ok_rows, "bad": bad[:12]}
    OUT["gather4"] = res
    print("gather4:", res)
    return all(v["status"] == 1 and v["rows_ok"] == 128 for v in res.values())


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def conv_case(m_in, m_out, cin, cout, K, seed, density=0.6, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    x = torch.zeros(m_in, ops.ceil4(cin))
    x[:, :cin] = torch.randn(m_in, cin, generator=g) * scale
    nbr = torch.randint(0, m_in, (m_out, K), generator=g, dtype=torch.int32)
    nbr[torch.rand(m_out, K, generator=g) > density] = -1
    W = torch.zeros(K, cin, ops.ceil4(cout))
    W[:, :, :cout] = torch.randn(K, cin, cout, generator=g) / (K * cin) ** 0.5
    bias = torch.randn(cout, generator=g)
    want = torch.zeros(m_out, cout, dtype=torch.float64)
    xd, Wd = x[:, :cin].double(), W[:, :, :cout].double()
    for k in range(K):
        ok = nbr[:, k] >= 0
        want[ok] += xd[nbr[ok, k].long()] @ Wd[k]
    want += bias.double()
    return x, nbr, W, bias, want.float()


def probe_conv(neg_mode):
    ops.HL_NEG_ROW_MODE = neg_mode
    res = []
    for (m_in, m_out, cin, cout, K) in [(300, 257, 16, 16, 27), (1000, 900, 80, 32, 27), (5000, 4100, 138, 16, 27),
                                        (700, 650, 32, 1, 27), (900, 300, 64, 64, 8), (3000, 2500, 24, 24, 27),
                                        (129, 129, 8, 8, 27), (640, 640, 40, 40, 27), (40000, 40000, 48, 24, 27),
                                        (2000, 2000, 192, 96, 27), (1500, 1500, 128, 128, 27), (700, 700, 160, 144, 27)]:
        x, nbr, W, bias, want = conv_case(m_in, m_out, cin, cout, K, seed=cin * 7 + cout)
        xc, Wc, bc, nc = x.cuda(), W.cuda(), bias.cuda(), nbr.cuda()
        ops.SPCONV_IMPL = "hl"
        try:
            y, part = ops.spconv(xc, cin, nc, Wc, cout, bias=bc, m_out=m_out, want_stats=True)
            torch.cuda.synchronize()
            err = rel(y[:, :cout].cpu(), want)
            s_ok = bool(torch.allclose(part.cpu()[:, 0].sum(0), y[:, :cout].cpu().sum(0), rtol=1e-4, atol=1e-3))
        except Exception as e:  # noqa: BLE001
            err, s_ok = repr(e)[:200] + f" debug_code={_lib.lib().ep_hl_debug_code()}", False
        res.append({"shape": (m_in, m_out, cin, cout, K), "rel_err": err, "bn_partial_ok": s_ok})
        print("conv", neg_mode, res[-1], flush=True)
    OUT[f"conv_neg{neg_mode}"] = res
    return all(isinstance(r["rel_err"], float) and r["rel_err"] < 3e-5 for r in res)


def morton(c):
    def spread(v):
        v = v.long() & 0x3FF
        v = (v | (v << 16)) & 0x30000FF
        v = (v | (v << 8)) & 0x300F00F
        v = (v | (v << 4)) & 0x30C30C3
        v = (v | (v << 2)) & 0x9249249
        return v
    return (spread(c[:, 0]) << 2) | (spread(c[:, 1]) << 1) | spread(c[:, 2])


def slab_set(n_side, thick=3):
    """two crossing slabs `thick` voxels thick in an n_side^3 grid, rows in Morton order (as the executor keeps them)"""
    ax = torch.arange(n_side)
    a = torch.stack(torch.meshgrid(ax, ax, torch.arange(thick) + n_side // 3, indexing="ij"), -1).view(-1, 3)
    b = torch.stack(torch.meshgrid(torch.arange(thick) + n_side // 2, ax, ax, indexing="ij"), -1).view(-1, 3)
    c = torch.unique(torch.cat([a, b]), dim=0)
    c = c[torch.argsort(morton(c))]
    return torch.cat([c, torch.zeros(c.shape[0], 1, dtype=c.dtype)], 1).int().contiguous()


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us


def probe_timing():
    from eprecon_b200.sparse import VoxelSet
    L = _lib.lib()
    res = []
    for n_side, shapes in ((182, [(74, 8), (48, 24), (24, 24), (32, 24)]), (72, [(138, 16), (96, 48), (48, 48), (64, 48)]),
                           (26, [(192, 96), (96, 96), (80, 32)]), (10, [(128, 128), (64, 64)])):
        coords = slab_set(n_side).cuda()
        vs = VoxelSet(coords, 1)
        nbr = vs.kmap_k3()
        m = coords.shape[0]
        pairs = int((nbr >= 0).sum().item())
        for cin, cout in shapes:
            g = torch.Generator().manual_seed(cin + cout)
            x = torch.zeros(m, ops.ceil4(cin))
            x[:, :cin] = torch.randn(m, cin, generator=g)
            W = torch.zeros(27, cin, ops.ceil4(cout))
            W[:, :, :cout] = torch.randn(27, cin, cout, generator=g) / (27 * cin) ** 0.5
            xc, Wc = x.cuda(), W.cuda()
            out = {}
            ys = {}
            for impl in ("tf32x3", "hl"):
                ops.SPCONV_IMPL = impl
                y, _ = ops.spconv(xc, cin, nbr, Wc, cout, m_out=m, want_stats=True)
                ys[impl] = y[:, :cout].clone()
                out[impl] = timeit(lambda: ops.spconv(xc, cin, nbr, Wc, cout, m_out=m, want_stats=True))
            # the conv launch alone (operand already split, as the producing epilogue will write it)
            w_hl, npad = ops._hl_weights(Wc, cout)
            x_hl = ops.hl_split(xc, cin)
            wsb = L.ep_spconv_hl_workspace_bytes(m, npad, 27)
            ws = torch.empty(max(wsb, 16), dtype=torch.uint8, device="cuda")
            o = torch.empty((m, ops.ceil4(cout)), dtype=torch.float32, device="cuda")
            part = torch.empty((L.ep_spconv_num_row_tiles(m), 2, cout), dtype=torch.float32, device="cuda")
            st = ops.stream_ptr()

            def conv_only():
                _lib.check(L.ep_spconv_hl_fwd(x_hl.data_ptr(), m, cin, nbr.data_ptr(), 27, w_hl.data_ptr(), npad, cout, 0,
                                              o.data_ptr(), o.stride(0), m, part.data_ptr(), ws.data_ptr(), wsb,
                                              ops.HL_NEG_ROW_MODE, st), "hl")
            out["hl_conv_only"] = timeit(conv_only)
            out["hl_split_only"] = timeit(lambda: ops.hl_split(xc, cin))
            r = {"m": m, "pairs": pairs, "cin": cin, "cout": cout, **{k: round(v, 1) for k, v in out.items()},
                 "hl_vs_tf32x3_rel": rel(ys["hl"], ys["tf32x3"]),
                 "gather_GBs_hl": round(pairs * ((cin + 31) // 32) * 128 / out["hl_conv_only"] / 1e3, 1)}
            res.append(r)
            print("timing", r, flush=True)
    OUT["timing"] = res


if __name__ == "__main__":
    # each part runs in its own process (a device trap poisons the CUDA context): --part gather4 | conv0 | conv1 | timing0 | timing1
    part = sys.argv[sys.argv.index("--part") + 1] if "--part" in sys.argv else "gather4"
    t0 = time.time()
    if part == "gather4":
        OUT["ok"] = probe_gather4()
    elif part.startswith("conv"):
        OUT["ok"] = probe_conv(int(part[4:]))
    elif part.startswith("timing"):
        ops.HL_NEG_ROW_MODE = int(part[6:])
        probe_timing()
    print(part, "ok:", OUT.get("ok"), "in %.1f s" % (time.time() - t0), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"probe_hl_{part}.json"), "w") as f:
        json.dump(OUT, f, indent=1, default=str)
