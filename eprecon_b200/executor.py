"""Python side of the native executor (csrc/executor.cu): parameter descriptors + scratch arenas + the four calls.

One ep_exec_* call replaces the ~160 (SPVCNN), ~60 (two ConvGRUs of a level), 6 (Linear4xTrans) or ~45 (occupancy
initialisation head) ctypes calls the Python mirrors in modules.py make, launching the same kernels in the same order
(bit-identical results, tests/test_executor_gpu.py).  ctypes releases the GIL for the whole call, so fragments on
different CUDA streams (eprecon_b200.streams.FragmentStreams) really run their host side in parallel.

Descriptor layout (flat int64, read sequentially by csrc/executor.cu::Reader):
  conv  = [W_ffma, W_umma_hi, W_umma_lo, W_hl, bias, K, cin, cout, npad] (pointers are 0 when unused; W_hl != 0 selects
          the TMA-gather kernel on half-pair operands, csrc/spconv_hl.cu)
  norm  = [gamma, beta, float32 bits of eps, channels]                  (BatchNorm1d or LayerNorm)
  res   = conv1, bn1, conv2, bn2, has_down, [conv_down, bn_down]
  SPVCNN      = cs[0..4], cin, stem conv+bn, down1 conv+bn, res, res, down2 conv+bn, res, res, pt0 lin+bn,
                dec1 conv+bn, res, res, dec2 conv+bn, res, res, pt1 lin+bn, END
  GRU level   = cv, ci, 2 x [convz (conv, lin), convr (conv, lin), convq (conv, lin)], END
  Linear4x    = n_heads, n_heads x [C_in, C_out, use_residual, lin1, ln1, lin2, ln2, lin3], END
  init head   = d, norm0, 7 x (conv, ln) of the ELAN, 3 x (conv, ln), subm4 conv, norm4, END
  decoder     = query_feat, query_embed, level_embed, gauss_B, decoder_norm (g, b), class_embed (w, b), mask_embed 3 x (w, b),
                6 x [ca in_proj (w, b), ca out_proj (w, b), ca norm (g, b), sa in_proj (w, b), sa out_proj (w, b), sa norm
                (g, b), ffn linear1 (w, b), linear2 (w, b), ffn norm (g, b)], END      (plain f32 pointers, csrc/decoder.cu)
  globals     = [k3 offsets stride 1, 2, 4; k2 offsets stride 1, 2; subm3 offsets]   (device int32 [K,3] tables)
Set EPRECON_EXEC=0 to run the per-kernel Python programs instead.
"""
import os
import struct
import threading

import numpy as np
import torch

from . import _lib, ops
from .ops import ceil4

ENABLED = os.environ.get("EPRECON_EXEC", "1") != "0"
ARENA_MB = int(os.environ.get("EPRECON_ARENA_MB", "1024"))
ARENA_MAX_MB = int(os.environ.get("EPRECON_ARENA_MAX_MB", "32768"))
_END = {"spvcnn": 0x5350564E, "gru": 0x47525546, "lin4x": 0x4C345854, "init": 0x494E4954, "decoder": 0x44454344}


def enabled():
    return ENABLED and ops.SPCONV_IMPL in ("tf32x3", "hl")


# ----------------------------------------------------------------------------------------------- descriptors
def _f2i(x):
    return struct.unpack("<i", struct.pack("<f", float(x)))[0]


class _Builder:
    def __init__(self):
        self.v, self.keep = [], []

    def ints(self, *xs):
        self.v.extend(int(x) for x in xs)

    def _conv(self, W, cout, bias):
        """W: FFMA-layout weights [K, cin, ceil4(cout)] (prepared by the module)."""
        K, cin = int(W.shape[0]), int(W.shape[1])
        hi = lo = hl = None
        npad = 0
        if K > 1:
            if ops.SPCONV_IMPL == "hl":
                hl, npad = ops._hl_weights(W, cout)
            else:
                hi, lo, npad = ops._umma_weights(W, cout, 3)
        self.keep += [W, hi, lo, hl, bias]
        self.v += [W.data_ptr(), hi.data_ptr() if hi is not None else 0, lo.data_ptr() if lo is not None else 0,
                   hl.data_ptr() if hl is not None else 0, bias.data_ptr() if bias is not None else 0, K, cin, int(cout),
                   int(npad)]

    def spconv(self, p):      # modules.SpConv3dParams (torchsparse conv: no bias)
        self._conv(p.prepared(), p.outc, None)

    def subm(self, p):        # modules.SubMConv3dParams (spconv SubMConv3d: bias)
        self._conv(p.prepared(), p.outc, p.bias.detach())

    def linear(self, lin, prep):
        self._conv(prep.get(), lin.weight.shape[0], lin.bias.detach() if lin.bias is not None else None)

    def norm(self, n):
        c = n.num_features if hasattr(n, "num_features") else n.normalized_shape[0]
        w, b = n.weight.detach(), n.bias.detach()
        self.keep += [w, b]
        self.v += [w.data_ptr(), b.data_ptr(), _f2i(n.eps), int(c)]

    def res(self, blk):       # modules.ResidualBlock
        net = blk.net
        self.spconv(net[0]); self.norm(net[1]); self.spconv(net[3]); self.norm(net[4])
        if len(blk.downsample) == 0:
            self.ints(0)
        else:
            self.ints(1)
            self.spconv(blk.downsample[0]); self.norm(blk.downsample[1])

    def finish(self, kind):
        self.ints(_END[kind])
        return np.asarray(self.v, dtype=np.int64), self.keep


def _build_spvcnn(m):
    b = _Builder()
    b.ints(*m.cs, m.in_channels)
    b.spconv(m.stem[0]); b.norm(m.stem[1])
    for stage in (m.stage1, m.stage2):
        b.spconv(stage[0].net[0]); b.norm(stage[0].net[1])
        b.res(stage[1]); b.res(stage[2])
    b.linear(m.point_transforms[0][0], m._pt[0]); b.norm(m.point_transforms[0][1])
    for up in (m.up1, m.up2):
        b.spconv(up[0].net[0]); b.norm(up[0].net[1])
        b.res(up[1][0]); b.res(up[1][1])
    b.linear(m.point_transforms[1][0], m._pt[1]); b.norm(m.point_transforms[1][1])
    return b.finish("spvcnn")


def _build_gru(pair):
    gv, gi = pair
    b = _Builder()
    b.ints(gv.hidden_dim, gi.hidden_dim)
    for g in (gv, gi):
        for s in (g.convz, g.convr, g.convq):
            b.spconv(s.net)
            b.linear(s.point_transforms[0], s._pl)
    return b.finish("gru")


def _build_lin4x(heads):
    b = _Builder()
    b.ints(len(heads))
    for h in heads:
        b.ints(h.C_in, h.C_out, int(h.use_residual))
        b.linear(h.linear1, h._p[0]); b.norm(h.norm1)
        b.linear(h.linear2, h._p[1]); b.norm(h.norm2)
        b.linear(h.linear3, h._p[2])
    return b.finish("lin4x")


def _build_init(m):
    b = _Builder()
    b.ints(m.dim)
    b.norm(m.norm0)
    e = m.similary_1
    for blk in (e.conv1, e.conv2, e.conv3, e.conv4, e.conv5, e.conv6, e.conv7):
        b.subm(blk.conv); b.norm(blk.ln)
    for conv, norm in ((m.subm1, m.norm1), (m.subm2, m.norm2), (m.subm3, m.norm3)):
        b.subm(conv.sparsesubmconv3d); b.norm(norm)
    b.subm(m.subm4.sparsesubmconv3d); b.norm(m.norm4)
    return b.finish("init")


def decoder_supported(m):
    """csrc/decoder.cu is compiled for the one configuration the reference instantiates (neucon_network.py:59-71)."""
    return (m.num_layers == 6 and m.num_queries == 80 and m.num_heads == 8 and m.query_feat.weight.shape[1] == 48
            and m.class_embed.weight.shape[0] == 21 and m.mask_embed.layers[0].weight.shape[0] == 192
            and m.mask_embed.layers[2].weight.shape[0] == 48 and m.transformer_ffn_layers[0].linear1.weight.shape[0] == 192
            and m.pos_enc.normalize and tuple(m.pos_enc.gauss_B.shape) == (3, 24))


def _build_decoder(m):
    b = _Builder()

    def ptrs(*ts):
        for t in ts:
            t = t.detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise _lib.EpreconError("decoder parameters must be contiguous float32")
            b.keep.append(t)
            b.v.append(t.data_ptr())
    ptrs(m.query_feat.weight, m.query_embed.weight, m.level_embed.weight, m.pos_enc.gauss_B,
         m.decoder_norm.weight, m.decoder_norm.bias, m.class_embed.weight, m.class_embed.bias)
    for lin in m.mask_embed.layers:
        ptrs(lin.weight, lin.bias)
    for j in range(m.num_layers):
        ca, sa = m.transformer_cross_attention_layers[j], m.transformer_self_attention_layers[j]
        ff = m.transformer_ffn_layers[j]
        ptrs(ca.multihead_attn.in_proj_weight, ca.multihead_attn.in_proj_bias, ca.multihead_attn.out_proj.weight,
             ca.multihead_attn.out_proj.bias, ca.norm.weight, ca.norm.bias,
             sa.self_attn.in_proj_weight, sa.self_attn.in_proj_bias, sa.self_attn.out_proj.weight, sa.self_attn.out_proj.bias,
             sa.norm.weight, sa.norm.bias, ff.linear1.weight, ff.linear1.bias, ff.linear2.weight, ff.linear2.bias,
             ff.norm.weight, ff.norm.bias)
    return b.finish("decoder")


def _params_of(obj):
    mods = obj if isinstance(obj, (tuple, list)) else (obj,)
    for m in mods:
        yield from m.parameters()


def _tensors_of(obj):
    for m in (obj if isinstance(obj, (tuple, list)) else (obj,)):
        yield from m.parameters()
        yield from m.buffers()


def _desc(owner, key, obj, build, tensors=_params_of):
    """Descriptor of `obj` (a module or a tuple of modules), cached on `owner` and rebuilt when a parameter tensor was
    moved (.to / .cuda) or modified in place (load_state_dict, optimizer step).  The Parameter objects themselves are
    looked up once: walking nn.Module trees on every call cost 3 ms per fragment."""
    cache = owner.__dict__.setdefault("_exec_desc", {})
    key = (key, ops.SPCONV_IMPL)
    hit = cache.get(key)
    if hit is not None:
        tag = tuple((p.data_ptr(), p._version) for p in hit[0])
        if tag == hit[1]:
            return hit[4]
    plist = list(tensors(obj))
    tag = tuple((p.data_ptr(), p._version) for p in plist)
    arr, keep = build(obj)
    cache[key] = (plist, tag, arr, keep, arr.ctypes.data)
    return arr.ctypes.data


_GLOBALS = {}


def _globals(device):
    key = str(device)
    g = _GLOBALS.get(key)
    if g is None:
        tabs = [ops.kernel_offsets("k3", s, device) for s in (1, 2, 4)] + \
               [ops.kernel_offsets("k2", s, device) for s in (1, 2)] + [ops.kernel_offsets("subm3", 1, device)]
        arr = np.asarray([t.data_ptr() for t in tabs], dtype=np.int64)
        g = _GLOBALS[key] = (arr, tabs, arr.ctypes.data)
    return g[2]


# --------------------------------------------------------------------------------------------------- arenas
_tls = threading.local()


def _state():
    st = getattr(_tls, "st", None)
    if st is None:
        st = _tls.st = {"arenas": {}, "stats": np.zeros(2, dtype=np.int64), "peak": 0}
    return st


def _run(fn, name, device, args_before, args_after=()):
    """fn(*args_before, arena_ptr, arena_bytes, stats_ptr, *args_after, stream); grows the arena on EP_ERR_WORKSPACE."""
    st = _state()
    stream = ops.stream_ptr()
    key = (str(device), stream)
    arena = st["arenas"].get(key)
    if arena is None:
        arena = st["arenas"][key] = torch.empty(ARENA_MB << 20, dtype=torch.uint8, device=device)
    stats = st["stats"]
    while True:
        status = fn(*args_before, arena.data_ptr(), arena.numel(), stats.ctypes.data, *args_after, stream)
        if status == -2 and arena.numel() < (ARENA_MAX_MB << 20):
            size = arena.numel() * 2
            st["arenas"][key] = arena = None      # release first: the caching allocator reuses it in stream order
            arena = st["arenas"][key] = torch.empty(size, dtype=torch.uint8, device=device)
            continue
        _lib.check(status, name)
        break
    _lib.LAUNCHES["n"] += int(stats[1])
    if stats[0] > st["peak"]:
        st["peak"] = int(stats[0])


def arena_peak_bytes():
    return _state()["peak"]


# ----------------------------------------------------------------------------------------------------- calls
def spvcnn(mod, feat, pts):
    """feat f32 [N, ld] (ld % 4 == 0, >= ceil4(in_channels), zero padded); pts f32 [N,4] contiguous -> [N, cs[4]]."""
    n = feat.shape[0]
    dev = feat.device
    cs4 = mod.cs[4]
    out = torch.empty((n, ceil4(cs4)), dtype=torch.float32, device=dev)
    _run(_lib.lib().ep_exec_spvcnn, "ep_exec_spvcnn", dev,
         (_desc(mod, "spvcnn", mod, _build_spvcnn), _globals(dev), pts.data_ptr(), feat.data_ptr(), feat.stride(0), n,
          float(mod.vres), out.data_ptr(), out.stride(0)))
    return out[:, :cs4]


def gru_level(owner, gru_v, gru_i, pts, gvalues, values, cv, c_all):
    """The voxel- and image-feature ConvGRUs of one level.  gvalues (hidden) / values (input): [u, ld] with columns
    [0,cv) voxel features and [cv,c_all) image features; returns the fused [u, c_all]."""
    u = values.shape[0]
    dev = values.device
    out = torch.empty((u, c_all), dtype=torch.float32, device=dev)
    _run(_lib.lib().ep_exec_gru_level, "ep_exec_gru_level", dev,
         (_desc(owner, ("gru", id(gru_v)), (gru_v, gru_i), _build_gru), _globals(dev), pts.data_ptr(), u, float(gru_v.vres),
          gvalues.data_ptr(), gvalues.stride(0), values.data_ptr(), values.stride(0), out.data_ptr(), out.stride(0)))
    return out


def linear4x(owner, heads, x):
    """heads: list of Linear4xTrans over the same input rows x [m, ld] -> list of [m, C_out] views."""
    m = x.shape[0]
    dev = x.device
    outs = [torch.empty((m, ceil4(h.C_out)), dtype=torch.float32, device=dev) for h in heads]
    ptrs = np.asarray([o.data_ptr() for o in outs], dtype=np.int64)
    _run(_lib.lib().ep_exec_linear4x, "ep_exec_linear4x", dev,
         (_desc(owner, ("lin4x",) + tuple(id(h) for h in heads), tuple(heads), _build_lin4x), x.data_ptr(), x.stride(0), m,
          ptrs.ctypes.data))
    return [o[:, :h.C_out] for o, h in zip(outs, heads)]


def init_head(mod, var, coords, shape):
    """var f32 [m, ld]; coords int32 [m,4]=(0,x,y,z); shape = site-grid dims -> logits [m,4] (column 0)."""
    m = var.shape[0]
    dev = var.device
    occ = torch.empty((m, 4), dtype=torch.float32, device=dev)
    _run(_lib.lib().ep_exec_init_head, "ep_exec_init_head", dev,
         (_desc(mod, "init", mod, _build_init), _globals(dev), var.data_ptr(), var.stride(0), coords.data_ptr(), m,
          int(shape[0]), int(shape[1]), int(shape[2]), occ.data_ptr()))
    return occ


def decoder(mod, rows, xyz, mask_rows, index, extent, want_aux=True):
    """MultiScaleMaskedTransformerDecoder.forward body (csrc/decoder.cu).  rows[l] f32 [N_l, ld] (48 channels); xyz[l] int64
    [N_l,3]; mask_rows f32 [N_2, ld]; index[0], index[1] int64 (nearest level-2 row of every level-0 / level-1 voxel)
    -> pred_logits [7,80,21] (prediction on query_feat + one per layer), pred_masks [80,N_2], aux_masks [6,80,N_2] | None."""
    L = _lib.lib()
    dev = mask_rows.device
    n = [int(r.shape[0]) for r in rows]
    logits = torch.empty((7, 80, 21), dtype=torch.float32, device=dev)
    masks = torch.empty((80, n[2]), dtype=torch.float32, device=dev)
    aux = torch.empty((6, 80, n[2]), dtype=torch.float32, device=dev) if want_aux else None
    ws_bytes = int(L.ep_exec_decoder_workspace_bytes(n[0], n[1], n[2]))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.check(L.ep_exec_decoder(_desc(mod, "decoder", mod, _build_decoder, _tensors_of), rows[0].data_ptr(), rows[0].stride(0),
                                 rows[1].data_ptr(), rows[1].stride(0), rows[2].data_ptr(), rows[2].stride(0),
                                 xyz[0].data_ptr(), xyz[1].data_ptr(), xyz[2].data_ptr(), n[0], n[1], n[2],
                                 mask_rows.data_ptr(), mask_rows.stride(0), index[0].data_ptr(), index[1].data_ptr(),
                                 float(extent[0]), float(extent[1]), float(extent[2]), logits.data_ptr(), masks.data_ptr(),
                                 aux.data_ptr() if want_aux else 0, ws.data_ptr(), ws_bytes, ops.stream_ptr()), "ep_exec_decoder")
    if want_aux:
        _lib.LAUNCHES["n"] += 4          # the full-level-2 mask passes of the aux predictions taken on levels 0 / 1
    return logits, masks, aux
