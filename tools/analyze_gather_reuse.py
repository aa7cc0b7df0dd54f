"""CPU analysis for the next kernel generation (no GPU needed): how much of the sparse-conv gather and of the
back-projection tap traffic is redundant inside one tile, i.e. what staging a tile's UNIQUE rows / feature-map sectors
in shared memory would save.  Uses the full-size synthetic fragment (level-2 candidates = children of the occupied 48^3
GT voxels, ~219 k voxels) and the kernels' own tiling: 128 Z-ordered rows per conv tile, 64 consecutive candidates per
back-projection tile.

    python tools/analyze_gather_reuse.py
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eprecon_b200 import synth  # noqa: E402
from oracle import restate  # noqa: E402


def spread3(v):
    v = v.astype(np.uint64) & np.uint64(0xFFFF)
    v = (v | (v << np.uint64(32))) & np.uint64(0x00FF00000000FFFF)
    v = (v | (v << np.uint64(16))) & np.uint64(0x00FF0000FF0000FF)
    v = (v | (v << np.uint64(8))) & np.uint64(0xF00F00F00F00F00F)
    v = (v | (v << np.uint64(4))) & np.uint64(0x30C30C30C30C30C3)
    v = (v | (v << np.uint64(2))) & np.uint64(0x9249249249249249)
    return v


def main():
    inputs, fa, fb = synth.make_fragment(seed=1)
    par = torch.nonzero(inputs["occ_list"][1][0]).numpy() * 2
    offs = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1]])
    xyz = (par[:, None, :] + offs[None]).reshape(-1, 3)
    n = len(xyz)
    # ---------------------------------------------------------------- sparse conv: 27-neighbourhood, Z-ordered 128-row tiles
    key = (spread3(xyz[:, 0] + 32768) << np.uint64(2)) | (spread3(xyz[:, 1] + 32768) << np.uint64(1)) | spread3(xyz[:, 2] + 32768)
    order = np.argsort(key, kind="stable")
    z = xyz[order]
    vol = np.full((100, 100, 100), -1, dtype=np.int64)
    vol[z[:, 0] + 1, z[:, 1] + 1, z[:, 2] + 1] = np.arange(n)
    nb = np.stack([vol[z[:, 0] + 1 + dx, z[:, 1] + 1 + dy, z[:, 2] + 1 + dz] for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)], 1)
    pairs = int((nb >= 0).sum())
    uniq, span = [], []
    for t0 in range(0, n, 128):
        r = nb[t0:t0 + 128]
        u = np.unique(r[r >= 0])
        uniq.append(len(u))
        span.append(int(u.max() - u.min() + 1))
    uniq, span = np.array(uniq), np.array(span)
    conv = {"voxels": n, "pairs_per_row": round(pairs / n, 2), "tiles": len(uniq),
            "gathered_rows_per_tile": round(pairs / len(uniq), 1), "unique_rows_per_tile_mean": round(float(uniq.mean()), 1),
            "unique_rows_per_tile_p95": int(np.percentile(uniq, 95)), "unique_rows_per_tile_max": int(uniq.max()),
            "gather_reduction_if_unique_rows_are_staged_once": round(pairs / float(uniq.sum()), 2),
            "row_index_span_per_tile_median": int(np.median(span)),
            "smem_bytes_for_unique_rows_p95": {f"cin{c}": int(np.percentile(uniq, 95)) * c * 4 for c in (16, 24, 48, 80)}}
    # ---------------------------------------------------------------- back-projection: 64-candidate tiles, 24 ch @ 120x160
    coords = torch.cat([torch.zeros(n, 1, dtype=torch.long), torch.from_numpy(xyz)], 1)
    kr = inputs["proj_matrices"][:, :, 0].permute(1, 0, 2, 3).contiguous()
    W, H, C = 160, 120, 24
    gx, gy, _, mask = restate.project_views(coords, inputs["vol_origin_partial"], 0.04, kr, H, W)
    with np.errstate(all="ignore"):
        ix = np.nan_to_num((gx + 1) / 2 * (W - 1), nan=0.0, posinf=0.0, neginf=0.0)
        iy = np.nan_to_num((gy + 1) / 2 * (H - 1), nan=0.0, posinf=0.0, neginf=0.0)
    x0, y0 = np.floor(ix).astype(np.int64), np.floor(iy).astype(np.int64)
    # tap reuse as a function of the tile size (consecutive candidates = children of consecutive parents in raster order)
    trend = {}
    for ts in (64, 256, 1024, 4096, 16384):
        tt, uu = 0, 0
        for t0 in range(0, n, ts):
            for v in range(mask.shape[0]):
                m = mask[v, t0:t0 + ts]
                if not m.any():
                    continue
                xs, ys = x0[v, t0:t0 + ts][m], y0[v, t0:t0 + ts][m]
                uu += len(np.unique(np.concatenate([(ys + dy) * (W + 2) + (xs + dx) for dx in (0, 1) for dy in (0, 1)])))
                tt += 4 * int(m.sum())
        trend[str(ts)] = {"tap_reuse": round(tt / uu, 2), "bytes_per_voxel_unique_pixels": round(uu * C * 4 / n, 1)}
    tot_taps, uniq_px, tiles = 0, 0, 0
    box_px = 0
    for t0 in range(0, n, 64):
        for v in range(mask.shape[0]):
            m = mask[v, t0:t0 + 64]
            if not m.any():
                continue
            xs, ys = x0[v, t0:t0 + 64][m], y0[v, t0:t0 + 64][m]
            px = np.unique(np.concatenate([(ys + dy) * (W + 2) + (xs + dx) for dx in (0, 1) for dy in (0, 1)]))
            tot_taps += 4 * int(m.sum())
            uniq_px += len(px)
            box_px += int((xs.max() - xs.min() + 2) * (ys.max() - ys.min() + 2))
        tiles += 1
    bp = {"tiles": tiles, "visible_voxel_views": tot_taps // 4, "taps": tot_taps, "unique_pixels_per_tile_view_total": uniq_px,
          "tap_reuse_inside_a_64_voxel_tile": round(tot_taps / uniq_px, 2),
          "bounding_box_pixels_over_unique_pixels": round(box_px / uniq_px, 2),
          "bytes_per_voxel_now": round(tot_taps * C * 4 / n, 1), "bytes_per_voxel_with_unique_pixels": round(uniq_px * C * 4 / n, 1),
          "bytes_per_voxel_with_bounding_boxes": round(box_px * C * 4 / n, 1),
          "algorithmic_bytes_per_voxel": round((4 * 9 * C * H * W + 20 * n + (16 + 4 * C) * n) / n, 1),
          "unique_pixel_traffic_by_tile_size": trend}
    print(json.dumps({"sparse_conv_level2": conv, "back_projection_level2": bp}, indent=1))


if __name__ == "__main__":
    main()
