"""Micro-benchmark of the sparse-conv kernels on realistic voxel sets (the three levels of the 96^3 bench fragment).
Prints one JSON line per (level, conv shape, implementation): avg launch time, algorithmic TFLOP/s and gather GB/s."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eprecon_b200 import ops, sparse, synth  # noqa: E402
from oracle import restate  # noqa: E402


def upsample8(c, interval):
    offs = torch.tensor([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1]]) * interval
    out = c.unsqueeze(1).repeat(1, 8, 1)
    out[:, :, 1:] += offs.unsqueeze(0)
    return out.view(-1, 4)


def main(iters=10, impls=("tf32x3", "ffma")):
    inputs, _, _ = synth.make_fragment(seed=1, with_features=False)
    dev = "cuda"
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    shapes = {0: [(80, 32, 27), (128, 128, 27), (192, 96, 27)], 1: [(138, 16, 27), (48, 48, 27), (96, 48, 27)],
              2: [(74, 8, 27), (24, 24, 27), (48, 24, 27), (24, 96, 1)]}
    for level in (2, 1, 0):
        scale = 2 - level
        interval = 2 ** scale
        vres = 0.04 * interval
        if level == 0:
            p = torch.nonzero(inputs["occ_list"][2][0]) * 4
            c = torch.cat([torch.zeros(len(p), 1, dtype=torch.long), p], 1)
        else:
            p = torch.nonzero(inputs["occ_list"][scale + 1][0]) * (interval * 2)
            c = upsample8(torch.cat([torch.zeros(len(p), 1, dtype=torch.long), p], 1), interval)
        pts = restate.aligned_points(c.int(), inputs["vol_origin_partial"], 0.04, inputs["world_to_aligned_camera"]).to(dev)
        pc = sparse.PointCloud(pts, vres)
        nbr = pc.vox.kmap_k3()
        m = pc.vox.m
        pairs = int((nbr >= 0).sum().item())
        for cin, cout, K in shapes[level]:
            x = torch.randn(m, ops.ceil4(cin), device=dev)
            W = torch.zeros(K, cin, ops.ceil4(cout), device=dev)
            W[:, :, :cout] = torch.randn(K, cin, cout, device=dev) / (K * cin) ** 0.5
            nb = nbr if K == 27 else None
            P = pairs if K == 27 else m
            for impl in impls:
                ops.SPCONV_IMPL = impl
                ops.spconv(x, cin, nb, W, cout, want_stats=True)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t = 0.0
                for _ in range(iters):
                    flush.zero_()
                    e0.record()
                    ops.spconv(x, cin, nb, W, cout, want_stats=True)
                    e1.record()
                    torch.cuda.synchronize()
                    t += e0.elapsed_time(e1) / iters
                flops = 2.0 * cin * cout * P
                gbytes = 4.0 * (P * cin + m * cout + K * cin * cout) + 4.0 * m * K
                print(json.dumps({"level": level, "m": m, "pairs_per_row": round(pairs / m, 2) if K == 27 else 1, "cin": cin,
                                  "cout": cout, "K": K, "impl": impl, "us": round(t * 1e3, 1),
                                  "TFLOPs": round(flops / t / 1e9, 2), "gather_GBs": round(gbytes / t / 1e6, 1)}))


if __name__ == "__main__":
    main()
