"""Thin tensor-level wrappers over the C ABI (libeprecon_b200.so).

Each wrapper validates dtype / contiguity / device, allocates outputs with torch (plumbing), passes
raw device pointers + the current CUDA stream, and raises on a non-zero status.  No CPU fallback.
"""
import os

import torch

from . import _lib

__all__ = ["stream_ptr", "to_nhwc", "backproject"]

# bench.py instrumentation: when PROFILE is a dict, the two dominant kernel families record a CUDA-event pair per
# launch on the launching stream ("events") or, in a separate untimed pass, their algorithmic work ("work").
PROFILE = None


def _prof_begin(kind):
    if PROFILE is None or PROFILE.get("mode") != "events":
        return None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    PROFILE.setdefault(kind, []).append((e0, e1))
    return e1


def _prof_end(e1):
    if e1 is not None:
        e1.record()


def _prof_work(kind, fn):
    if PROFILE is not None and PROFILE.get("mode") == "work":
        PROFILE.setdefault(kind + "_work", []).append(fn())


_DEV_INDEX = None


def stream_ptr():
    """Raw cudaStream_t of torch's current stream.  torch.cuda.current_stream() costs ~13 us of Python per call (it was
    7 ms of an 800-launch step); the raw query is ~0.3 us.  One process drives one GPU, so the device index is cached."""
    global _DEV_INDEX
    if _DEV_INDEX is None:
        _DEV_INDEX = torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(_DEV_INDEX)


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


def pack_views_nhwc(views):
    """Per-view feature maps -> ONE channels-last buffer [V,bs,H,W,C] (the layout the gather kernels consume).
    `views`: list of V tensors [bs,C,H,W] (the reference's per-view backbone outputs, neuralrecon.py:53-54) -> one
    ep_pack_views_nhwc launch (no torch.stack copy, no separate transpose); or a single tensor [V,bs,C,H,W] / an already
    channels-last [V,bs,H,W,C]-strided one -> to_nhwc (zero-copy when the caller stores channels-last)."""
    import ctypes
    if torch.is_tensor(views):
        return to_nhwc(views.float())
    v0 = views[0]
    bs, C, H, W = v0.shape
    if not all(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (bs, C, H, W) for t in views):
        return to_nhwc(torch.stack([t.float() for t in views]))
    V = len(views)
    if V > 32:
        return to_nhwc(torch.stack(list(views)))
    out = torch.empty((V, bs, H, W, C), dtype=torch.float32, device=v0.device)
    ptrs = (ctypes.c_void_p * V)(*[t.data_ptr() for t in views])
    _lib.check(_lib.lib().ep_pack_views_nhwc(ptrs, V, bs, C, H * W, out.data_ptr(), stream_ptr()), "ep_pack_views_nhwc")
    return out


def backproject(coords, origin, voxel_size, feats_nhwc, krcam, min_views, mode="mean", out=None, out_col=0,
                want_src=False, want_zbar=False, min_valid=1, alloc_width=None):
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
    mode_i = {"mean": 0, "meanvar": 1}[mode]
    count = torch.empty(n, dtype=torch.float32, device=dev)
    counters = torch.empty(bs + 1, dtype=torch.int32, device=dev)
    if BP_IMPL == "fused" and C in (24, 32, 40, 80) and out is None:
        # single pass: outputs are allocated at capacity n and sliced to the survivor count afterwards
        width = alloc_width or C
        buf = torch.empty((n, width), dtype=torch.float32, device=dev)
        out_coords = torch.empty((n, 4), dtype=torch.int32, device=dev)
        out_vis = torch.empty(n, dtype=torch.int32, device=dev)
        src = torch.empty(n, dtype=torch.int32, device=dev) if want_src else None
        zbar = torch.empty(n, dtype=torch.float32, device=dev) if want_zbar else None
        wsb = L.ep_backproject_fused_workspace_bytes(n)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        _prof_work("bp_gather", lambda: {"bytes": 0, "n_in": n, "C": C, "deferred": True})
        _e = _prof_begin("bp_gather")
        _lib.check(L.ep_backproject_fused(coords.data_ptr(), n, origin.data_ptr(), float(voxel_size), krcam.data_ptr(), V, bs,
                                          H, W, feats_nhwc.data_ptr(), C, int(min_views), mode_i, count.data_ptr(),
                                          out_coords.data_ptr(), out_vis.data_ptr(), _ptr(src), buf.data_ptr(), width,
                                          _ptr(zbar), counters.data_ptr(), ws.data_ptr(), wsb, st), "ep_backproject_fused")
        _prof_end(_e)
        host = counters.tolist()   # the one host sync of this op (output size is data dependent)
        m = host[bs]
        if PROFILE is not None and PROFILE.get("mode") == "work":
            PROFILE["bp_gather_work"][-1] = {"bytes": 4 * V * C * H * W + 64 * V + 20 * n + (16 + 4 * C) * m, "n_in": n,
                                             "n_out": m, "C": C}
        if any(v < min_valid for v in host[:bs]):
            return None
        return {"feat": buf[:m, :C], "coords": out_coords[:m], "count": count, "vis": out_vis[:m],
                "src": src[:m] if src is not None else None, "zbar": zbar[:m] if zbar is not None else None,
                "n_valid": host[:bs], "buffer": buf[:m]}
    vis = torch.empty(n, dtype=torch.int32, device=dev)
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
        # alloc_width: allocate a wider row (the caller concatenates more columns after the C sampled ones)
        out = torch.empty((m, alloc_width or C), dtype=torch.float32, device=dev)
        out_col = 0
    else:
        _chk(out, torch.float32, "out")
        assert out.shape[0] == m and out_col % 4 == 0 and out_col + C <= out.shape[1]
    zbar = torch.empty(m, dtype=torch.float32, device=dev) if want_zbar else None
    # algorithmic bytes of the whole back-projection call: maps + KRt + coords in + count out + rows out
    _prof_work("bp_gather", lambda: {"bytes": 4 * V * C * H * W + 64 * V + 20 * n + (16 + 4 * C) * m, "n_in": n,
                                     "n_out": m, "C": C})
    _e = _prof_begin("bp_gather")
    _lib.check(L.ep_backproject_gather(out_coords.data_ptr(), out_vis.data_ptr(), m, feats_nhwc.data_ptr(), C, V, bs,
                                       H, W, origin.data_ptr(), float(voxel_size), krcam.data_ptr(), mode_i,
                                       out.data_ptr() + 4 * out_col, out.shape[1], _ptr(zbar), st), "ep_backproject_gather")
    _prof_end(_e)
    return {"feat": out[:, out_col:out_col + C], "coords": out_coords, "count": count, "vis": out_vis, "src": src,
            "zbar": zbar, "n_valid": host[:bs], "buffer": out}


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


# =====================================================================================================
# generic plumbing
# =====================================================================================================
U64_MAX = 0xFFFFFFFFFFFFFFFF


def ceil4(c):
    return (c + 3) // 4 * 4


def _L():
    return _lib.lib()


def compact_flags(flags, want_pos=False):
    """uint8/bool flags [n] -> (index int32 [total] ascending, total) [+ pos int32 [n]]; one host sync."""
    n = flags.shape[0]
    dev = flags.device
    flags = flags.view(torch.uint8) if flags.dtype == torch.bool else flags
    L = _L()
    wsb = L.ep_compact_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    index = torch.empty(n, dtype=torch.int32, device=dev)
    pos = torch.empty(n, dtype=torch.int32, device=dev) if want_pos else None
    total = torch.empty(1, dtype=torch.int32, device=dev)
    _lib.check(L.ep_compact_flags(flags.data_ptr(), n, index.data_ptr(), _ptr(pos), total.data_ptr(), ws.data_ptr(),
                                  wsb, stream_ptr()), "ep_compact_flags")
    t = int(total.item())
    return (index[:t], t, pos) if want_pos else (index[:t], t)


def sort_segments(keys, key_bits, sentinel=U64_MAX, want_seg_of_item=True):
    """Group items by 64-bit key (int64 storage).  Returns dict(keys_sorted, perm, seg_start, seg_end, seg_of_item, S)."""
    n = keys.shape[0]
    dev = keys.device
    L = _L()
    wsb = L.ep_sort_segments_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    ks = torch.empty(n, dtype=torch.int64, device=dev)
    perm = torch.empty(n, dtype=torch.int32, device=dev)
    seg_start = torch.empty(n, dtype=torch.int32, device=dev)
    seg_end = torch.empty(n, dtype=torch.int32, device=dev)
    soi = torch.empty(n, dtype=torch.int32, device=dev) if want_seg_of_item else None
    nseg = torch.empty(1, dtype=torch.int32, device=dev)
    _lib.check(L.ep_sort_segments(keys.data_ptr(), n, int(key_bits), sentinel, ks.data_ptr(), perm.data_ptr(),
                                  seg_start.data_ptr(), seg_end.data_ptr(), _ptr(soi), nseg.data_ptr(), ws.data_ptr(),
                                  wsb, stream_ptr()), "ep_sort_segments")
    S = int(nseg.item())
    return {"keys_sorted": ks, "perm": perm, "seg_start": seg_start[:S], "seg_end": seg_end[:S], "seg_of_item": soi,
            "S": S}


class HashTable:
    """Open-addressing table key(int64 bits) -> row id, load factor <= 0.5."""

    def __init__(self, keys):
        m = keys.shape[0]
        cap = 1 << max(4, (2 * m - 1).bit_length())
        self.cap = cap
        self.keys = torch.empty(cap, dtype=torch.int64, device=keys.device)
        self.vals = torch.empty(cap, dtype=torch.int32, device=keys.device)
        _lib.check(_L().ep_hash_build(keys.data_ptr(), m, self.keys.data_ptr(), self.vals.data_ptr(), cap,
                                      stream_ptr()), "ep_hash_build")


def coord_keys(coords, batch_first):
    m = coords.shape[0]
    keys = torch.empty(m, dtype=torch.int64, device=coords.device)
    _lib.check(_L().ep_coord_keys(coords.data_ptr(), m, int(batch_first), keys.data_ptr(), stream_ptr()),
               "ep_coord_keys")
    return keys


_OFFSET_CACHE = {}


def kernel_offsets(kind, stride, device):
    """int32 [K,3] offset tables: 'k3' (torchsparse odd volume: z-outer/x-inner), 'k2' (even: x-outer/z-inner),
    'subm3' (spconv weight order: x-outer/z-inner)."""
    key = (kind, stride, str(device))
    if key not in _OFFSET_CACHE:
        if kind == "k3":
            o = [[x, y, z] for z in (-1, 0, 1) for y in (-1, 0, 1) for x in (-1, 0, 1)]
        elif kind == "k2":
            o = [[x, y, z] for x in (0, 1) for y in (0, 1) for z in (0, 1)]
        elif kind == "subm3":
            o = [[x, y, z] for x in (-1, 0, 1) for y in (-1, 0, 1) for z in (-1, 0, 1)]
        else:
            raise ValueError(kind)
        _OFFSET_CACHE[key] = (torch.tensor(o, dtype=torch.int32) * stride).to(device)
    return _OFFSET_CACHE[key]


def kmap_build(out_coords, batch_first, offsets, table, shape=None):
    m, K = out_coords.shape[0], offsets.shape[0]
    nbr = torch.empty((m, K), dtype=torch.int32, device=out_coords.device)
    sx, sy, sz = shape if shape is not None else (0, 0, 0)
    _lib.check(_L().ep_kmap_build(out_coords.data_ptr(), m, int(batch_first), offsets.data_ptr(), K,
                                  table.keys.data_ptr(), table.vals.data_ptr(), table.cap, int(sx), int(sy), int(sz),
                                  nbr.data_ptr(), stream_ptr()), "ep_kmap_build")
    return nbr


# =====================================================================================================
# dense-per-row math
# =====================================================================================================
# back-projection: "fused" = single-pass kernel (count + look-back compaction + gather), "3pass" = count / compact / gather
BP_IMPL = os.environ.get("EPRECON_BP", "fused")

# "hl" (default) = tcgen05 kind::f16 on pre-split half-pair operands gathered by cp.async (csrc/spconv_hl.cu); "tf32x3" / "tf32" =
# the round-1 tcgen05 kind::tf32 kernel with register producers (csrc/spconv_tc.cu); "ffma" = fp32 CUDA-core gather-GEMM (csrc/spconv.cu)
SPCONV_IMPL = os.environ.get("EPRECON_SPCONV", "hl")
_UMMA_CACHE = {}


def _umma_weights(W, cout, prec):
    """FFMA-layout weights [K, cin, ceil4(cout)] -> (w_hi, w_lo, npad) in the UMMA slab layout [K][nq][npad][4], nq = 4 * ceil(cin / 16)."""
    key = (W.data_ptr(), W._version, tuple(W.shape), prec)
    hit = _UMMA_CACHE.get(key)
    if hit is None:
        K, cin, _ = W.shape
        npad = (cout + 15) // 16 * 16
        if npad > 128:
            npad = (cout + 127) // 128 * 128
        nq = (cin + 15) // 16 * 4          # quads of input channels, zero-padded to whole 16-channel slabs
        wp = torch.zeros((K, nq * 4, npad), dtype=torch.float32, device=W.device)
        wp[:, :cin, :cout] = W[:, :, :cout]
        u = wp.view(torch.int32)
        hi = ((u + 0xFFF + ((u >> 13) & 1)) & ~0x1FFF).view(torch.float32)     # round-to-nearest-even tf32
        lo = None
        if prec == 3:
            lo = (wp - hi).view(K, nq, 4, npad).permute(0, 1, 3, 2).contiguous()  # exact remainder
        hi = hi.view(K, nq, 4, npad).permute(0, 1, 3, 2).contiguous()
        if len(_UMMA_CACHE) > 4096:
            _UMMA_CACHE.clear()
        hit = _UMMA_CACHE[key] = (hi, lo, npad, W)   # keep W alive so the data_ptr key stays unique
    return hit[0], hit[1], hit[2]


_HL_CACHE = {}
HL_NEG_ROW_MODE = int(os.environ.get("EPRECON_HL_NEG", "0"))


def _hl_weights(W, cout):
    """FFMA-layout weights [K, cin, ceil4(cout)] -> (w_hl uint16-as-half [K][nslab][npad][64], npad): per 32-channel slab
    and output channel the 32 halfs h = fp16(w) followed by the 32 halfs fp16((w - h) * 2^11) (csrc/spconv_hl.cu)."""
    key = (W.data_ptr(), W._version, tuple(W.shape))
    hit = _HL_CACHE.get(key)
    if hit is None:
        K, cin, _ = W.shape
        npad = (cout + 15) // 16 * 16
        if npad > 128:
            npad = (cout + 127) // 128 * 128
        nslab = (cin + 31) // 32
        wp = torch.zeros((K, nslab * 32, npad), dtype=torch.float32, device=W.device)
        wp[:, :cin, :cout] = W[:, :, :cout]
        h = wp.half()
        lo = ((wp - h.float()) * 2048.0).half()
        h4 = h.view(K, nslab, 32, npad).permute(0, 1, 3, 2)
        l4 = lo.view(K, nslab, 32, npad).permute(0, 1, 3, 2)
        w_hl = torch.cat([h4, l4], dim=3).contiguous()
        if len(_HL_CACHE) > 4096:
            _HL_CACHE.clear()
        hit = _HL_CACHE[key] = (w_hl, npad, W)
    return hit[0], hit[1]


_COUNTERS = {}


def _counters(dev):
    """Zeroed int32 ticket buffer of the CURRENT stream for the fused split-K reduction of ep_spconv_hl_fused_fwd (the
    kernels leave it zeroed; one buffer per stream so concurrent streams never share tickets)."""
    key = (str(dev), stream_ptr())
    t = _COUNTERS.get(key)
    if t is None:
        t = _COUNTERS[key] = torch.zeros(1024, dtype=torch.int32, device=dev)
    return t


def hl_split(x, c, overflow=None):
    """fp32 rows [m, ld] -> half-pair rows [m, nslab*64] (fp16 storage) for ep_spconv_hl_fwd."""
    m = x.shape[0]
    nslab = (c + 31) // 32
    out = torch.empty((m, nslab * 64), dtype=torch.float16, device=x.device)
    _lib.check(_L().ep_hl_split_rows(x.data_ptr(), x.stride(0), c, m, out.data_ptr(), _ptr(overflow), stream_ptr()),
               "ep_hl_split_rows")
    return out


def spconv(x, cin, nbr, W, cout, bias=None, m_out=None, want_stats=False, out=None, out_col=0):
    """out[j,:cout] = bias + sum_k W[k]^T x[nbr[j,k], :cin].  x [*, ld]; W [K, cin, ceil4(cout)] prepared by the
    module.  Returns (out [m_out, ceil4(cout)] or view into `out`, bn_partial or None)."""
    L = _L()
    K = W.shape[0]
    if m_out is None:
        m_out = nbr.shape[0] if nbr is not None else x.shape[0]
    dev = x.device
    if out is None:
        out = torch.empty((m_out, ceil4(cout)), dtype=torch.float32, device=dev)
        if ceil4(cout) != cout:
            out[:, cout:].zero_()
        out_col = 0
    part = None
    if want_stats:
        part = torch.empty((L.ep_spconv_num_row_tiles(m_out), 2, cout), dtype=torch.float32, device=dev)
    _prof_work("spconv", lambda: {"pairs": int((nbr >= 0).sum().item()) if nbr is not None else int(m_out),
                                  "cin": cin, "cout": cout, "K": K, "m_out": int(m_out), "m_in": int(x.shape[0])})
    _e = _prof_begin("spconv")
    if SPCONV_IMPL == "ffma" or K == 1:   # dense per-row linears (2 slabs) stay on the CUDA-core GEMM: no gather to hide
        _lib.check(L.ep_spconv_fwd(x.data_ptr(), x.stride(0), cin, _ptr(nbr), K, W.data_ptr(), W.shape[2], cout,
                                   _ptr(bias), out.data_ptr() + 4 * out_col, out.stride(0), m_out, _ptr(part),
                                   stream_ptr()), "ep_spconv_fwd")
    elif SPCONV_IMPL == "hl":
        w_hl, npad = _hl_weights(W, cout)
        x_hl = hl_split(x, cin)
        wsb = L.ep_spconv_hl_workspace_bytes(m_out, npad, K)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev) if wsb else None
        ctr = _counters(dev)
        _lib.check(L.ep_spconv_hl_fused_fwd(x_hl.data_ptr(), x.shape[0], cin, _ptr(nbr), K, w_hl.data_ptr(), npad, cout,
                                            _ptr(bias), out.data_ptr() + 4 * out_col, out.stride(0), m_out, _ptr(part),
                                            _ptr(ws), wsb, HL_NEG_ROW_MODE, ctr.data_ptr(), ctr.numel(), 0, 0, 0.0, 0,
                                            stream_ptr()), "ep_spconv_hl_fused_fwd")
    else:
        prec = 3 if SPCONV_IMPL == "tf32x3" else 1
        w_hi, w_lo, npad = _umma_weights(W, cout, prec)
        wsb = L.ep_spconv_tc_workspace_bytes(m_out, npad, K)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev) if wsb else None
        _lib.check(L.ep_spconv_tc_fwd(x.data_ptr(), x.stride(0), cin, _ptr(nbr), K, w_hi.data_ptr(), _ptr(w_lo), npad,
                                      cout, _ptr(bias), out.data_ptr() + 4 * out_col, out.stride(0), m_out, _ptr(part),
                                      prec, _ptr(ws), wsb, stream_ptr()), "ep_spconv_tc_fwd")
    _prof_end(_e)
    return out, part


def colstats(x, c):
    m = x.shape[0]
    L = _L()
    part = torch.empty((L.ep_spconv_num_row_tiles(m), 2, c), dtype=torch.float32, device=x.device)
    _lib.check(L.ep_colstats(x.data_ptr(), x.stride(0), m, c, part.data_ptr(), stream_ptr()), "ep_colstats")
    return part


def bn_scale_shift(part, m, gamma, beta, eps=1e-5):
    c = part.shape[2]
    ss = torch.empty((2, c), dtype=torch.float32, device=part.device)
    _lib.check(_L().ep_bn_finalize(part.data_ptr(), part.shape[0], c, m, float(eps), _ptr(gamma), _ptr(beta),
                                   ss.data_ptr(), 0, stream_ptr()), "ep_bn_finalize")
    return ss


def affine_act(a, c, ss_a=None, b=None, ss_b=None, relu=False, out=None, out_col=0):
    m = a.shape[0]
    if out is None:
        out, out_col = a, 0
    _lib.check(_L().ep_affine_act(a.data_ptr(), a.stride(0), _ptr(ss_a), _ptr(b), b.stride(0) if b is not None else 0,
                                  _ptr(ss_b), int(relu), m, c, out.data_ptr() + 4 * out_col, out.stride(0),
                                  stream_ptr()), "ep_affine_act")
    return out


def layernorm(x, c, gamma, beta, res=None, relu_before=False, relu_after=False, eps=1e-5, out=None):
    m = x.shape[0]
    if out is None:
        out = x
    _lib.check(_L().ep_layernorm(x.data_ptr(), x.stride(0), _ptr(res), res.stride(0) if res is not None else 0,
                                 int(relu_before), _ptr(gamma), _ptr(beta), float(eps), int(relu_after), m, c,
                                 out.data_ptr(), out.stride(0), stream_ptr()), "ep_layernorm")
    return out


def gather_rows(src, c, index=None, shift=0, fill=0.0, m=None, out=None, out_col=0, src_col=0):
    if m is None:
        m = index.shape[0] if index is not None else src.shape[0]
    if out is None:
        out = torch.empty((m, ceil4(c)), dtype=torch.float32, device=src.device)
        if ceil4(c) != c:
            out[:, c:].zero_()
        out_col = 0
    if m > 0:
        _lib.check(_L().ep_gather_rows(src.data_ptr() + 4 * src_col, src.stride(0), _ptr(index), int(shift),
                                       float(fill), m, c, out.data_ptr() + 4 * out_col, out.stride(0), stream_ptr()),
                   "ep_gather_rows")
    return out


def gather_coords(src, index):
    m = index.shape[0]
    out = torch.empty((m, 4), dtype=torch.int32, device=src.device)
    if m > 0:
        _lib.check(_L().ep_gather_coords(src.data_ptr(), index.data_ptr(), m, out.data_ptr(), stream_ptr()),
                   "ep_gather_coords")
    return out


def aligned_coords(coords, origin, voxel_size, w2ac, zero_batch=False):
    n = coords.shape[0]
    out = torch.empty((n, 4), dtype=torch.float32, device=coords.device)
    _lib.check(_L().ep_aligned_coords(coords.data_ptr(), n, origin.data_ptr(), float(voxel_size), w2ac.data_ptr(),
                                      int(zero_batch), out.data_ptr(), stream_ptr()), "ep_aligned_coords")
    return out


def upsample8(coords, interval):
    n = coords.shape[0]
    out = torch.empty((n * 8, 4), dtype=torch.int32, device=coords.device)
    _lib.check(_L().ep_upsample8(coords.data_ptr(), n, int(interval), out.data_ptr(), stream_ptr()), "ep_upsample8")
    return out


def threshold_flags(x, thr, mode=0):
    n = x.shape[0]
    flags = torch.empty(n, dtype=torch.uint8, device=x.device)
    _lib.check(_L().ep_threshold_flags(x.data_ptr(), x.stride(0), n, float(thr), int(mode), flags.data_ptr(),
                                       stream_ptr()), "ep_threshold_flags")
    return flags


def masked_attention(q, k, v, blocked, n_heads, scale):
    """Fused masked cross-attention (csrc/attention.cu): q [Q, H*6], k / v [N, H*6] projected rows, blocked bool / uint8
    [Q, N] (True = may not attend) or None -> [Q, H*6].  One pass over the keys, deterministic."""
    L = _L()
    q = _chk(q, torch.float32, "q")
    k = _chk(k, torch.float32, "k")
    v = _chk(v, torch.float32, "v")
    nq, e = q.shape
    n = k.shape[0]
    if blocked is not None:
        blocked = blocked.view(torch.uint8) if blocked.dtype == torch.bool else blocked
        blocked = _chk(blocked, torch.uint8, "blocked")
        assert tuple(blocked.shape) == (nq, n)
    out = torch.empty((nq, e), dtype=torch.float32, device=q.device)
    wsb = L.ep_masked_attention_workspace_bytes(n, n_heads)
    ws = torch.empty(wsb, dtype=torch.uint8, device=q.device)
    _lib.check(L.ep_masked_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), k.stride(0), _ptr(blocked), n, nq, n_heads,
                                     e // n_heads, float(scale), out.data_ptr(), ws.data_ptr(), wsb, stream_ptr()),
               "ep_masked_attention")
    return out

