"""Micro-benchmark of the back-projection kernels at the three levels of the 96^3 config.

Times count / compact / gather separately with CUDA events (L2 flushed between iterations) and prints
achieved algorithmic GB/s per SURVEY.md section 8(d):  B = 4VCHW + 64V + 20 N_in + (16+4C) N_out.
"""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eprecon_b200 import _lib, ops, synth  # noqa: E402


def upsample8(c, interval):
    offs = torch.tensor([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1]],
                        dtype=c.dtype) * interval
    out = c.unsqueeze(1).repeat(1, 8, 1)
    out[:, :, 1:] += offs.unsqueeze(0)
    return out.view(-1, 4)


def main(iters=20):
    dev = "cuda"
    inputs, fa, fb = synth.make_fragment(seed=1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    L = _lib.lib()
    for level in (0, 1, 2):
        scale = 2 - level
        interval = 2 ** scale
        if level == 0:
            g = torch.stack(torch.meshgrid(*[torch.arange(0, 96, 4)] * 3, indexing="ij")).view(3, -1)
            coords = torch.cat([torch.zeros(1, g.shape[1], dtype=torch.long), g]).t().contiguous().int()
            minv = 2
        else:
            occ = inputs["occ_list"][scale + 1][0]
            p = torch.nonzero(occ) * (interval * 2)
            p = torch.cat([torch.zeros(len(p), 1, dtype=torch.long), p], 1)
            coords = upsample8(p, interval).int().contiguous()
            minv = 0
        feats = torch.stack([f[scale] for f in fb]).to(dev)
        kr = inputs["proj_matrices"][:, :, scale].permute(1, 0, 2, 3).contiguous().to(dev)
        origin = inputs["vol_origin_partial"].to(dev)
        coords = coords.to(dev)
        V, bs, C, H, W = feats.shape
        nhwc = ops.to_nhwc(feats)
        ops.BP_IMPL = "3pass"
        res = ops.backproject(coords, origin, 0.04, nhwc, kr, minv, want_src=True)
        ops.BP_IMPL = "fused"
        n, m = coords.shape[0], res["coords"].shape[0]
        out = torch.empty((m, C), device=dev)
        count = torch.empty(n, device=dev)
        vis = torch.empty(n, dtype=torch.int32, device=dev)
        counters = torch.empty(2, dtype=torch.int32, device=dev)
        wsb = L.ep_backproject_workspace_bytes(n)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        oc = torch.empty((m, 4), dtype=torch.int32, device=dev)
        ov = torch.empty(m, dtype=torch.int32, device=dev)
        st = ops.stream_ptr()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        t = [0.0] * 4
        for it in range(iters + 3):
            flush.zero_()
            ev[0].record()
            ops.to_nhwc(feats)
            ev[1].record()
            L.ep_backproject_count(coords.data_ptr(), n, origin.data_ptr(), 0.04, kr.data_ptr(), V, bs, H, W, minv,
                                   count.data_ptr(), vis.data_ptr(), counters.data_ptr(), counters[1:].data_ptr(),
                                   ws.data_ptr(), wsb, st)
            ev[2].record()
            L.ep_backproject_compact(coords.data_ptr(), vis.data_ptr(), n, minv, oc.data_ptr(), ov.data_ptr(), 0,
                                     ws.data_ptr(), st)
            ev[3].record()
            L.ep_backproject_gather(oc.data_ptr(), ov.data_ptr(), m, nhwc.data_ptr(), C, V, bs, H, W,
                                    origin.data_ptr(), 0.04, kr.data_ptr(), 0, out.data_ptr(), C, 0, st)
            ev[4].record()
            torch.cuda.synchronize()
            if it >= 3:
                for k in range(4):
                    t[k] += ev[k].elapsed_time(ev[k + 1]) / iters
        # fused single-pass kernel (same outputs, one launch)
        bufF = torch.empty((n, C), device=dev)
        ocF = torch.empty((n, 4), dtype=torch.int32, device=dev)
        ovF = torch.empty(n, dtype=torch.int32, device=dev)
        totF = torch.empty(2, dtype=torch.int32, device=dev)
        wsf = L.ep_backproject_fused_workspace_bytes(n)
        wsF = torch.empty(wsf, dtype=torch.uint8, device=dev)
        tf = 0.0
        ef = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for it in range(iters + 3):
            flush.zero_()
            ef[0].record()
            L.ep_backproject_fused(coords.data_ptr(), n, origin.data_ptr(), 0.04, kr.data_ptr(), V, bs, H, W, nhwc.data_ptr(), C,
                                   minv, 0, count.data_ptr(), ocF.data_ptr(), ovF.data_ptr(), 0, bufF.data_ptr(), C, 0,
                                   totF.data_ptr(), wsF.data_ptr(), wsf, st)
            ef[1].record()
            torch.cuda.synchronize()
            if it >= 3:
                tf += ef[0].elapsed_time(ef[1]) / iters
        assert int(totF[1].item()) == m and torch.equal(ocF[:m], oc) and torch.allclose(bufF[:m], out, rtol=1e-5, atol=1e-6)
        visible = float(res["count"].sum().item()) / max(n, 1)
        bytes_alg = 4 * V * C * H * W + 64 * V + 20 * n + (16 + 4 * C) * m
        tot = t[1] + t[2] + t[3]
        print(json.dumps({"level": level, "n_in": n, "n_out": m, "C": C, "HW": [H, W], "mean_visible_views": round(visible, 2),
                          "ms_transpose": round(t[0], 4), "ms_count": round(t[1], 4), "ms_compact": round(t[2], 4),
                          "ms_gather": round(t[3], 4), "ms_fused_single_pass": round(tf, 4),
                          "GBs_fused": round(bytes_alg / tf / 1e6, 1), "frac_fused": round(bytes_alg / tf / 1e6 / peak, 3), "alg_MB": round(bytes_alg / 1e6, 2),
                          "GBs_all3": round(bytes_alg / tot / 1e6, 1), "frac_all3": round(bytes_alg / tot / 1e6 / peak, 3),
                          "GBs_gather_only": round((4 * V * C * H * W + (20 + 4 * C) * m) / t[3] / 1e6, 1)}))


def batched(B):
    """One launch over B fragments' level-2 candidates (operands >> L2): the HBM-bound view (same probe as bench.py)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    inputs, _, _ = synth.make_fragment(seed=1)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    except Exception:
        pass
    print(json.dumps(bench.bp_batched_probe(torch.device("cuda", 0), inputs, peaks, B=B)))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--batched":
        for b in sys.argv[2].split(","):
            batched(int(b))
    else:
        main()
