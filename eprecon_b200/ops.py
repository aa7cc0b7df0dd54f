"""Thin tensor-level wrappers over the C ABI (libeprecon_b200.so).

Each wrapper validates dtype / contiguity / device, allocates outputs with torch (plumbing), passes
raw device pointers + the current CUDA stream, and raises on a non-zero status.  No CPU fallback.
"""
import torch

from . import _lib

__all__ = ["stream_ptr", "to_nhwc", "backproject"]


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, dtype, name):
    if not t.is_cuda:
        raise _lib.EpreconError(f"{name} must be a CUDA tensor (eprecon_b200 has no CPU path)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    return t


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def to_nhwc(feats):
    """[V,bs,C,H,W] (any strides) -> channels-last buffer [V,bs,H,W,C]; zero-copy when the input is
    already stored channels-last, otherwise one tiled transpose kernel."""
    V, bs, C, H, W = feats.shape
    cl = feats.permute(0, 1, 3, 4, 2)
    if cl.is_contiguous():
        return cl
    src = _chk(feats.contiguous(), torch.float32, "feats")
    out = torch.empty((V, bs, H, W, C), dtype=torch.float32, device=feats.device)
    _lib.check(_lib.lib().ep_nchw_to_nhwc(src.data_ptr(), out.data_ptr(), V * bs, C, H * W, stream_ptr()),
               "ep_nchw_to_nhwc")
    return out


def backproject(coords, origin, voxel_size, feats_nhwc, krcam, min_views, mode="mean", out=None, out_col=0,
                want_src=False, want_zbar=False, min_valid=1):
    """Project + visibility + stable compaction + bilinear gather.

    coords int32 [N,4] (b,x,y,z); origin f32 [bs,3]; feats_nhwc f32 [V,bs,H,W,C]; krcam f32 [V,bs,4,4].
    Returns None when any batch entry has fewer than `min_valid` surviving voxels (the reference's
    degenerate-fragment convention), else a dict with feat [M,C] (or a view into `out`), coords [M,4],
    count [N], vis [M] (bit v = visible in view v), src [M] (row of `coords` each survivor came from).
    """
    L = _lib.lib()
    coords = _chk(coords, torch.int32, "coords")
    origin = _chk(origin, torch.float32, "origin")
    feats_nhwc = _chk(feats_nhwc, torch.float32, "feats_nhwc")
    krcam = _chk(krcam, torch.float32, "krcam")
    V, bs, H, W, C = feats_nhwc.shape
    n = coords.shape[0]
    dev = coords.device
    st = stream_ptr()
    count = torch.empty(n, dtype=torch.float32, device=dev)
    vis = torch.empty(n, dtype=torch.int32, device=dev)
    counters = torch.empty(bs + 1, dtype=torch.int32, device=dev)
    ws_bytes = L.ep_backproject_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.check(L.ep_backproject_count(coords.data_ptr(), n, origin.data_ptr(), float(voxel_size), krcam.data_ptr(),
                                      V, bs, H, W, int(min_views), count.data_ptr(), vis.data_ptr(),
                                      counters.data_ptr(), counters[bs:].data_ptr(), ws.data_ptr(), ws_bytes, st),
               "ep_backproject_count")
    host = counters.tolist()  # the one host sync of this op (output size is data dependent)
    m = host[bs]
    if any(v < min_valid for v in host[:bs]):
        return None
    out_coords = torch.empty((m, 4), dtype=torch.int32, device=dev)
    out_vis = torch.empty(m, dtype=torch.int32, device=dev)
    src = torch.empty(m, dtype=torch.int32, device=dev) if want_src else None
    _lib.check(L.ep_backproject_compact(coords.data_ptr(), vis.data_ptr(), n, int(min_views), out_coords.data_ptr(),
                                        out_vis.data_ptr(), _ptr(src), ws.data_ptr(), st), "ep_backproject_compact")
    if out is None:
        out = torch.empty((m, C), dtype=torch.float32, device=dev)
        out_col = 0
    else:
        _chk(out, torch.float32, "out")
        assert out.shape[0] == m and out_col % 4 == 0 and out_col + C <= out.shape[1]
    zbar = torch.empty(m, dtype=torch.float32, device=dev) if want_zbar else None
    _lib.check(L.ep_backproject_gather(out_coords.data_ptr(), out_vis.data_ptr(), m, feats_nhwc.data_ptr(), C, V, bs,
                                       H, W, origin.data_ptr(), float(voxel_size), krcam.data_ptr(),
                                       {"mean": 0, "meanvar": 1}[mode], out.data_ptr() + 4 * out_col, out.shape[1],
                                       _ptr(zbar), st), "ep_backproject_gather")
    return {"feat": out[:, out_col:out_col + C], "coords": out_coords, "count": count, "vis": out_vis, "src": src,
            "zbar": zbar, "n_valid": host[:bs]}


def backproject_grid(res, origin, voxel_size, krcam, V, bs, H, W):
    """Materialise the reference's im_grid [V,M,2] and mask [V,M] (dead downstream, kept for the 5-tuple)."""
    m = res["coords"].shape[0]
    dev = res["coords"].device
    im_grid = torch.empty((V, m, 2), dtype=torch.float32, device=dev)
    mask = torch.empty((V, m), dtype=torch.bool, device=dev)
    _lib.check(_lib.lib().ep_backproject_grid(res["coords"].data_ptr(), res["vis"].data_ptr(), m, V, bs, H, W,
                                              origin.data_ptr(), float(voxel_size), krcam.data_ptr(),
                                              im_grid.data_ptr(), mask.data_ptr(), stream_ptr()),
               "ep_backproject_grid")
    return im_grid, mask
