"""Panoptic decoder — drop-in for models/mask3dformer.py (MultiScaleMaskedTransformerDecoder.forward :338-429,
forward_prediction_heads :431-447, panoptic_post / panoptic_inference :462-581) and the Fourier voxel position encoding
(models/voxel_position_encoding.py:42-70,116-146).  SURVEY.md section 8(f) row 1: first widening after the TSDF path.

Same constructor keywords, forward() signature, output dict and state-dict names as the reference
(`query_feat`, `query_embed`, `transformer_{self,cross}_attention_layers.N.{self_attn,multihead_attn}.in_proj_*`,
`...out_proj`, `transformer_ffn_layers.N.linear{1,2}`, `decoder_norm`, `level_embed`, `class_embed`,
`mask_embed.layers.N`, `pos_enc.gauss_B`).

What changes underneath:
  * the reference's nearest-voxel search materialises `cdist(level-2 voxels, level-l voxels)` (105 k x 26 k floats =
    11 GB at the 96^3 config) and arg-mins it; here the level-2 voxels go into the coordinate hash table
    (csrc/hash.cu) and every coarse voxel probes the box that must contain its nearest level-2 voxel
    (`ep_kmap_build`), then takes the minimum of (integer distance, row) -- identical result incl. the first-index
    tie rule, 21 MB instead of 11 GB;
  * attention masks are never expanded to [heads, Q, N] booleans: one [Q, N] mask is shared by the heads;
  * panoptic_inference does its per-query bookkeeping with three bincounts and ONE device->host read instead of
    three `.item()` syncs per query.
  * the masked cross-attention over the voxel keys (6.2 of the decoder's 10 ms on the library route: batched GEMMs with
    an inner dimension of 6) is one fused pass over the keys, `ep_masked_attention` (csrc/attention.cu).
  * the whole forward (K/V projections with the position encoding folded in, cross- and self-attention, FFN, prediction
    heads, attention masks) is ONE native call, `ep_exec_decoder` (csrc/decoder.cu, 44 launches + 4 when the aux
    predictions are requested), instead of 419 launches issued from Python in round 1.  EPRECON_NATIVE_DECODER=0 keeps
    the per-op formulation (fused cross-attention + cuBLAS/ATen for the small GEMMs) that the native call is tested
    against (tests/test_zz_panoptic_decoder_gpu.py).
CUDA tensors only: there is no CPU path.
"""
import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import executor, ops
from ._lib import EpreconError


# cross-attention over the voxel keys through the fused kernel (csrc/attention.cu); the ATen formulation remains for the
# 80 x 80 self-attention and for decoder shapes the kernel does not cover (head_dim != 6)
FUSED_ATTENTION = os.environ.get("EPRECON_FUSED_ATTENTION", "1") != "0"
NATIVE_DECODER = os.environ.get("EPRECON_NATIVE_DECODER", "1") != "0"


def _need_cuda(t, what):
    if not t.is_cuda:
        raise EpreconError(f"{what}: expected a CUDA tensor (eprecon_b200 has no CPU path)")


class PositionEmbeddingCoordsSine(nn.Module):
    """Fourier-feature position encoding (`pos_type='fourier'`, `normalize=True`): xyz [B,N,3] -> [B,d_pos,N]."""

    def __init__(self, temperature=10000, normalize=False, scale=None, pos_type="fourier", d_pos=None, d_in=3, gauss_scale=1.0):
        super().__init__()
        if pos_type != "fourier":
            raise NotImplementedError("the reference decoder only instantiates pos_type='fourier' (mask3dformer.py:257)")
        assert d_pos is not None and d_pos % 2 == 0
        self.d_pos, self.normalize = d_pos, normalize
        self.register_buffer("gauss_B", torch.empty((d_in, d_pos // 2)).normal_() * gauss_scale)

    def rows(self, xyz, extent):
        """xyz [N,3] (any integer / float dtype), extent (sx,sy,sz) -> [N, d_pos] rows (sin | cos)."""
        x = xyz.to(torch.float32)
        if self.normalize:
            x = x / torch.tensor([float(s) for s in extent], dtype=torch.float32, device=x.device)
        proj = (x * (2 * math.pi)) @ self.gauss_B
        return torch.cat([proj.sin(), proj.cos()], 1)

    def forward(self, xyz, num_channels=None, input_range=None):
        assert xyz.ndim == 3
        extent = (input_range[1] - input_range[0]).reshape(-1)[:3].tolist() if input_range is not None else (1.0, 1.0, 1.0)
        return torch.stack([self.rows(x, extent).t() for x in xyz])


class MLP(nn.Module):
    """Linear stack with ReLU between the layers (state-dict names `layers.N.{weight,bias}` as in the reference)."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        widths = [input_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
        self.num_layers = num_layers
        self.layers = nn.ModuleList([nn.Linear(widths[i], widths[i + 1]) for i in range(num_layers)])

    def forward(self, x):
        last = self.num_layers - 1
        for i in range(self.num_layers):
            x = self.layers[i](x)
            x = x if i == last else F.relu(x)
        return x


class _Attention(nn.Module):
    """Parameter holder with nn.MultiheadAttention's names; `attend` works on 2-D [rows, E] operands (batch of one)."""

    def __init__(self, d_model, nhead):
        super().__init__()
        self.embed_dim, self.num_heads = d_model, nhead
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d_model, d_model))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d_model))
        self.out_proj = nn.Linear(d_model, d_model)
        nn.init.xavier_uniform_(self.in_proj_weight)

    def project_memory(self, key_rows, value_rows):
        """K, V of a memory level as [heads, N, head_dim]."""
        E, h = self.embed_dim, self.num_heads
        k = F.linear(key_rows, self.in_proj_weight[E:2 * E], self.in_proj_bias[E:2 * E])
        v = F.linear(value_rows, self.in_proj_weight[2 * E:], self.in_proj_bias[2 * E:])
        return k.view(-1, h, E // h).transpose(0, 1), v.view(-1, h, E // h).transpose(0, 1)

    def attend(self, query_rows, key_rows, value_rows, blocked=None, fused=False):
        E, h = self.embed_dim, self.num_heads
        if fused and FUSED_ATTENTION and E // h == 6 and h <= 8 and query_rows.shape[0] <= 96:
            # cross-attention over the voxel keys: one fused pass (score -> mask -> online softmax -> weighted sum,
            # csrc/attention.cu) instead of [heads, Q, N] score tensors and two batched GEMMs with a dimension of 6
            q = F.linear(query_rows, self.in_proj_weight[:E], self.in_proj_bias[:E])
            k = F.linear(key_rows, self.in_proj_weight[E:2 * E], self.in_proj_bias[E:2 * E])
            v = F.linear(value_rows, self.in_proj_weight[2 * E:], self.in_proj_bias[2 * E:])
            return self.out_proj(ops.masked_attention(q, k, v, blocked, h, 1.0 / math.sqrt(E // h)))
        q = F.linear(query_rows, self.in_proj_weight[:E], self.in_proj_bias[:E]).view(-1, h, E // h).transpose(0, 1)
        k, v = self.project_memory(key_rows, value_rows)
        s = torch.matmul(q, k.transpose(1, 2)) * (1.0 / math.sqrt(E // h))       # [h, Q, N]
        if blocked is not None:
            s.masked_fill_(blocked.unsqueeze(0), float("-inf"))                   # one [Q, N] mask for all heads
        a = torch.softmax(s, dim=-1)
        o = torch.matmul(a, v).transpose(0, 1).reshape(-1, E)
        return self.out_proj(o)


class SelfAttentionLayer(nn.Module):
    def __init__(self, d_model, nhead, dropout=0.0, activation="relu", normalize_before=False):
        super().__init__()
        if normalize_before or dropout != 0.0:
            raise NotImplementedError("the reference builds the decoder with pre_norm=False, dropout=0 (neucon_network.py:62-71)")
        self.self_attn = _Attention(d_model, nhead)
        self.norm = nn.LayerNorm(d_model)

    def forward(self, tgt, tgt_mask=None, tgt_key_padding_mask=None, query_pos=None):
        qk = tgt if query_pos is None else tgt + query_pos
        return self.norm(tgt + self.self_attn.attend(qk, qk, tgt))


class CrossAttentionLayer(nn.Module):
    def __init__(self, d_model, nhead, dropout=0.0, activation="relu", normalize_before=False):
        super().__init__()
        if normalize_before or dropout != 0.0:
            raise NotImplementedError("the reference builds the decoder with pre_norm=False, dropout=0 (neucon_network.py:62-71)")
        self.multihead_attn = _Attention(d_model, nhead)
        self.norm = nn.LayerNorm(d_model)

    def forward(self, tgt, memory, memory_mask=None, memory_key_padding_mask=None, pos=None, query_pos=None):
        q = tgt if query_pos is None else tgt + query_pos
        k = memory if pos is None else memory + pos
        return self.norm(tgt + self.multihead_attn.attend(q, k, memory, memory_mask, fused=True))


class FFNLayer(nn.Module):
    def __init__(self, d_model, dim_feedforward=2048, dropout=0.0, activation="relu", normalize_before=False):
        super().__init__()
        if normalize_before or dropout != 0.0 or activation != "relu":
            raise NotImplementedError("post-norm / relu / no dropout is the only configuration the reference uses")
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm = nn.LayerNorm(d_model)

    def forward(self, tgt):
        return self.norm(tgt + self.linear2(F.relu(self.linear1(tgt))))


_BOX_CACHE = {}


def _box_offsets(radius, device):
    key = (radius, str(device))
    if key not in _BOX_CACHE:
        r = torch.arange(-radius, radius + 1, dtype=torch.int32)
        o = torch.stack(torch.meshgrid(r, r, r, indexing="ij"), -1).view(-1, 3).contiguous()
        _BOX_CACHE[key] = (o.to(device), (o.long() ** 2).sum(1).to(device))
    return _BOX_CACHE[key]


def nearest_fine_index(coarse, fine, radius):
    """For every coarse voxel the row of its nearest voxel of `fine` (Euclidean, first row on ties) -- what
    `argmin(cdist(fine, coarse[None]), dim=1)` returns (mask3dformer.py:361-368).  coarse / fine: int32 [N,4] (b,x,y,z).
    `radius` must bound the per-axis distance of the true nearest neighbour: after the level alignment every level-1
    voxel contains a level-2 voxel (distance^2 <= 3 -> radius 1) and every level-0 voxel one within its 4^3 block
    (distance^2 <= 27 -> radius 5).  Hash join (ep_hash_build + ep_kmap_build) instead of the N x M distance matrix."""
    _need_cuda(coarse, "nearest_fine_index")
    table = ops.HashTable(ops.coord_keys(fine, True))
    offs, d2 = _box_offsets(radius, coarse.device)
    nbr = ops.kmap_build(coarse, True, offs, table)                               # [Nc, K] rows of `fine`, -1 = absent
    big = 1 << 24
    key = torch.where(nbr >= 0, d2.unsqueeze(0) * big + nbr.long(), torch.full((), 1 << 62, dtype=torch.int64, device=nbr.device))
    best = key.min(dim=1).values
    if bool((best >= (1 << 62)).any()):
        raise EpreconError("nearest_fine_index: a coarse voxel has no level-2 voxel within its search box "
                           "(the levels were not aligned, neucon_network.py:516-541)")
    return best % big


class MultiScaleMaskedTransformerDecoder(nn.Module):
    def __init__(self, mask_classification=True, *, num_classes, hidden_dim, num_queries, nheads, dim_feedforward, dec_layers,
                 pre_norm, mask_dim):
        super().__init__()
        assert mask_classification, "Only support mask classification model"
        self.mask_classification = mask_classification
        self.pos_enc_type = "fourier"
        self.num_queries, self.num_heads, self.num_layers = num_queries, nheads, dec_layers
        self.query_feat = nn.Embedding(num_queries, hidden_dim)
        self.query_embed = nn.Embedding(num_queries, hidden_dim)
        self.transformer_self_attention_layers = nn.ModuleList()
        self.transformer_cross_attention_layers = nn.ModuleList()
        self.transformer_ffn_layers = nn.ModuleList()
        self.pos_enc = PositionEmbeddingCoordsSine(pos_type="fourier", d_pos=mask_dim, gauss_scale=1.0, normalize=True)
        for _ in range(dec_layers):
            self.transformer_self_attention_layers.append(SelfAttentionLayer(hidden_dim, nheads, normalize_before=pre_norm))
            self.transformer_cross_attention_layers.append(CrossAttentionLayer(hidden_dim, nheads, normalize_before=pre_norm))
            self.transformer_ffn_layers.append(FFNLayer(hidden_dim, dim_feedforward, normalize_before=pre_norm))
        self.decoder_norm = nn.LayerNorm(hidden_dim)
        self.num_feature_levels = 3
        self.level_embed = nn.Embedding(self.num_feature_levels, hidden_dim)
        self.class_embed = nn.Linear(hidden_dim, num_classes + 1)
        self.mask_embed = MLP(hidden_dim, hidden_dim * 4, mask_dim, 3)
        # the reference always returns the six intermediate predictions; only training consumes them -- callers that read
        # pred_logits / pred_masks alone (NeuConNet's inference path) clear this and save 4 passes over [80, N_2]
        self.aux_outputs = True

    def forward_prediction_heads(self, output, mask_rows, index):
        """output [Q,E]; mask_rows [N2,E]; index int64 [N_l] (None = all level-2 voxels) -> class logits [Q,classes+1],
        mask logits [Q,N2], blocked [Q,N_l] (True = the query may not attend to that voxel)."""
        d = self.decoder_norm(output)
        logits = self.class_embed(d)
        masks = self.mask_embed(d) @ mask_rows.t()
        sel = masks if index is None else masks.index_select(1, index)
        return logits, masks, sel < 0            # sigmoid(x) < 0.5  <=>  x < 0

    @torch.no_grad()
    def forward(self, panoptic_features, panoptic_coords, mask_features, spitial_shape):
        """panoptic_features[l] [1,C,N_l]; panoptic_coords[l] [1,N_l,3] (xyz, 4 cm units); mask_features [1,C,N_2]
        -> {'pred_logits' [1,Q,classes+1], 'pred_masks' [1,Q,N_2], 'aux_outputs': [...]} (the reference's dict)."""
        _need_cuda(mask_features, "MultiScaleMaskedTransformerDecoder.forward")
        if mask_features.shape[0] != 1:
            raise EpreconError("the decoder is called per fragment (batch of one), as in neucon_network.py:560-575")
        rows = [f[0].t().contiguous() for f in panoptic_features]                     # [N_l, C]
        xyz = [c[0] for c in panoptic_coords]
        mask_rows = mask_features[0].t().contiguous()
        c4 = [torch.cat([torch.zeros_like(x[:, :1]), x], 1).to(torch.int32).contiguous() for x in xyz]
        index = [nearest_fine_index(c4[0], c4[2], 5), nearest_fine_index(c4[1], c4[2], 1), None]
        if NATIVE_DECODER and min(r.shape[0] for r in rows) > 0:
            if not executor.decoder_supported(self):
                raise EpreconError("the native decoder is built for the reference configuration (48 channels, 8 heads, 80 "
                                   "queries, 6 layers, FFN 192, 20 classes); set EPRECON_NATIVE_DECODER=0 for other shapes")
            extent = [float(s) for s in (spitial_shape.reshape(-1)[:3].tolist() if torch.is_tensor(spitial_shape) else spitial_shape)]
            logits, masks, aux = executor.decoder(self, rows, [x.to(torch.int64).contiguous() for x in xyz], mask_rows, index,
                                                  extent, want_aux=self.aux_outputs)
            out = {"pred_logits": logits[6:7], "pred_masks": masks.unsqueeze(0)}
            if self.aux_outputs:
                out["aux_outputs"] = [{"pred_logits": logits[j:j + 1], "pred_masks": aux[j:j + 1]} for j in range(6)]
            return out
        src, keys = [], []
        for l in range(3):
            s = rows[l] + self.level_embed.weight[l].unsqueeze(0)
            src.append(s)
            keys.append(s + self.pos_enc.rows(xyz[l], spitial_shape))
        qpos = self.query_embed.weight
        out = self.query_feat.weight
        logits, masks, blocked = self.forward_prediction_heads(out, mask_rows, index[0])
        aux = [(logits, masks)]
        for j in range(self.num_layers):
            l = j % self.num_feature_levels
            blocked = blocked & ~blocked.all(dim=1, keepdim=True)                     # fully blocked query -> attends everywhere
            ca = self.transformer_cross_attention_layers[j]
            out = ca.norm(out + ca.multihead_attn.attend(out + qpos, keys[l], src[l], blocked, fused=True))
            out = self.transformer_self_attention_layers[j](out, query_pos=qpos)
            out = self.transformer_ffn_layers[j](out)
            logits, masks, blocked = self.forward_prediction_heads(out, mask_rows, index[(j + 1) % self.num_feature_levels])
            aux.append((logits, masks))
        return {"pred_logits": logits.unsqueeze(0), "pred_masks": masks.unsqueeze(0),
                "aux_outputs": [{"pred_logits": a.unsqueeze(0), "pred_masks": b.unsqueeze(0)} for a, b in aux[:-1]]}


def panoptic_inference(mask_cls, mask_pred, object_mask_threshold=0.3, thing_id=list(range(3, 21)), overlap_threshold=0.5):
    """mask3dformer.py:516-581 with one host read: mask_cls [Q,classes+1], mask_pred [Q,N] -> [panoptic_seg int32 [N], info]."""
    _need_cuda(mask_pred, "panoptic_inference")
    Q, N = mask_pred.shape
    scores, labels = F.softmax(mask_cls, dim=-1).max(-1)
    keep = labels.ne(0) & (scores > object_mask_threshold)
    prob = mask_pred.sigmoid()
    weighted = torch.where(keep.unsqueeze(1), scores.unsqueeze(1) * prob, torch.full((), -1.0, device=prob.device))
    winner = weighted.argmax(0)                                                        # first query on ties, as torch.argmax
    confident = prob.gather(0, winner.unsqueeze(0)).squeeze(0) >= 0.5                  # the winner's own mask covers the voxel
    area = torch.bincount(winner, minlength=Q)
    inter = torch.bincount(winner[confident], minlength=Q)
    orig = (prob >= 0.5).sum(1)
    host = torch.stack([keep.long(), labels, area, orig, inter]).cpu().tolist()       # the one device->host read
    seg_of_query = [0] * Q
    info, stuff, next_id = [], {}, 0
    for q in range(Q):
        if not host[0][q]:
            continue
        cls, a, o, m = host[1][q], host[2][q], host[3][q], host[4][q]
        if a > 0 and o > 0 and m > 0:
            if a / o < overlap_threshold:
                continue
            thing = cls in thing_id
            if not thing:
                if cls in stuff:
                    seg_of_query[q] = stuff[cls]
                    continue
                stuff[cls] = next_id + 1
            next_id += 1
            seg_of_query[q] = next_id
            info.append({"id": next_id, "isthing": bool(thing), "category_id": int(cls)})
    table = torch.tensor(seg_of_query, dtype=torch.int32, device=mask_pred.device)
    seg = torch.where(confident, table[winner], torch.zeros((), dtype=torch.int32, device=mask_pred.device))
    if not any(host[0]):
        seg = torch.zeros(N, dtype=torch.int32, device=mask_pred.device)
    return [seg, info]


def panoptic_post(outputs, semantic_on=False, panoptic_on=True, instance_on=False, occupied=None):
    """mask3dformer.py:461-500 (panoptic branch; semantic / instance inference are off at every reference call site)."""
    if semantic_on or instance_on:
        raise NotImplementedError("only panoptic_on is used by the reference (neucon_network.py:586)")
    cls = outputs["pred_logits"]
    msk = outputs["pred_masks"] if occupied is None else outputs["pred_masks"][..., occupied]
    res = {}
    for c, m in zip(cls, msk):
        res["panoptic_seg"] = panoptic_inference(c, m)
    return res
