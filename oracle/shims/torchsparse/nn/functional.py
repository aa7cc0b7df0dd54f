"""Functional primitives of torchsparse v2.0.0, restated in numpy / PyTorch for CPU.

Semantics restated from the published torchsparse v2.0.0 sources (hash_cuda.cu, hashmap,
count/voxelize/devoxelize kernels, nn/functional/{conv,downsample}.py).  Reference call sites:
ops/torchsparse_utils.py:19-27,44-58,71-100 and every `spnn.Conv3d` in models/modules.py.
"""
import numpy as np
import torch

from ..tensor import SparseTensor
from .utils import get_kernel_offsets

__all__ = ["sphash", "sphashquery", "spcount", "spvoxelize", "spdevoxelize", "calc_ti_weights",
           "spdownsample", "conv3d"]

_FNV_OFFSET = np.uint64(14695981039346656037)
_FNV_PRIME = np.uint64(1099511628211)
_MASK60 = np.uint64(0x0FFFFFFFFFFFFFFF)


def _hash_np(c):
    """c: int32 ndarray [..., 4] (x,y,z,b) -> uint64 60-bit FNV-1a-style hash."""
    with np.errstate(over="ignore"):
        h = np.full(c.shape[:-1], _FNV_OFFSET, dtype=np.uint64)
        for j in range(4):
            h = h ^ c[..., j].astype(np.uint32).astype(np.uint64)
            h = h * _FNV_PRIME
        h = (h >> np.uint64(60)) ^ (h & _MASK60)
    return h


def sphash(coords, offsets=None):
    """coords int[N,4]; offsets int[K,3] or None -> int64 [N] or [K,N] (offset added to xyz only)."""
    c = coords.detach().cpu().numpy().astype(np.int32)
    if offsets is None:
        return torch.from_numpy(_hash_np(c).astype(np.int64))
    o = offsets.detach().cpu().numpy().astype(np.int32)
    cc = np.repeat(c[None, :, :], o.shape[0], axis=0)  # [K,N,4]
    cc[:, :, :3] += o[:, None, :]
    return torch.from_numpy(_hash_np(cc).astype(np.int64))


def sphashquery(queries, references):
    """Position of each query hash in `references` (first occurrence), -1 when absent."""
    q = queries.detach().cpu().numpy().astype(np.int64)
    r = references.detach().cpu().numpy().astype(np.int64)
    if r.size == 0:
        return torch.full(q.shape, -1, dtype=torch.int64)
    order = np.argsort(r, kind="stable")
    rs = r[order]
    pos = np.searchsorted(rs, q.reshape(-1), side="left")
    pos_c = np.minimum(pos, rs.size - 1)
    hit = rs[pos_c] == q.reshape(-1)
    out = np.where(hit, order[pos_c], -1).astype(np.int64)
    return torch.from_numpy(out.reshape(q.shape))


def spcount(idx, num):
    idx = idx.long()
    valid = idx >= 0
    return torch.bincount(idx[valid], minlength=num).int()


def spvoxelize(feats, idx, counts):
    """Mean-pool rows of `feats` into voxels: out[idx[i]] += feats[i] / counts[idx[i]]."""
    idx = idx.long()
    valid = idx >= 0
    out = torch.zeros(counts.shape[0], feats.shape[1], dtype=feats.dtype)
    if valid.any():
        iv = idx[valid]
        out.index_add_(0, iv, feats[valid] / counts[iv].to(feats.dtype).unsqueeze(1))
    return out


def spdevoxelize(feats, idx, weights):
    """feats [M,C]; idx [N,8]; weights [N,8] -> [N,C] = sum_k w * feats[idx] over idx>=0."""
    idx = idx.long()
    out = torch.zeros(idx.shape[0], feats.shape[1], dtype=feats.dtype)
    for k in range(idx.shape[1]):
        ik = idx[:, k]
        ok = ik >= 0
        if ok.any():
            out[ok] += weights[ok, k].unsqueeze(1) * feats[ik[ok]]
    return out


def calc_ti_weights(coords, idx_query, scale=1):
    """Trilinear weights [8,N], corner order x-outer/z-inner, zeroed on misses, renormalised."""
    with torch.no_grad():
        p = coords
        pf = torch.floor(coords / scale) * scale if scale != 1 else torch.floor(coords)
        pc = pf + scale
        x, y, z = (p[:, i].view(-1, 1) for i in range(3))
        xf, yf, zf = (pf[:, i].view(-1, 1).float() for i in range(3))
        xc, yc, zc = (pc[:, i].view(-1, 1).float() for i in range(3))
        w = torch.cat([
            (xc - x) * (yc - y) * (zc - z), (xc - x) * (yc - y) * (z - zf),
            (xc - x) * (y - yf) * (zc - z), (xc - x) * (y - yf) * (z - zf),
            (x - xf) * (yc - y) * (zc - z), (x - xf) * (yc - y) * (z - zf),
            (x - xf) * (y - yf) * (zc - z), (x - xf) * (y - yf) * (z - zf)], dim=1)
        w = w.transpose(1, 0).contiguous()
        if scale != 1:
            w /= scale ** 3
        w[idx_query == -1] = 0
        w /= torch.sum(w, dim=0) + 1e-8
    return w


def _t(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v, v)


def spdownsample(coords, stride=2, kernel_size=2, tensor_stride=1):
    """Output sites of a strided conv whose kernel tiles the stride (k == s): truncate each
    coordinate toward zero to a multiple of stride*tensor_stride, then unique sorted by (b,x,y,z)."""
    stride, kernel_size, tensor_stride = _t(stride), _t(kernel_size), _t(tensor_stride)
    assert all(stride[k] in (1, kernel_size[k]) for k in range(3)), "only k==s strided convs are on the hot path"
    ss = torch.tensor([stride[k] * tensor_stride[k] for k in range(3)], dtype=torch.int).unsqueeze(0)
    coords = coords.clone()
    coords[:, :3] = (torch.div(coords[:, :3], ss.float()).trunc() * ss).to(coords.dtype)
    coords = coords[:, [3, 0, 1, 2]]
    coords = torch.unique(coords, dim=0)
    return coords[:, [1, 2, 3, 0]].contiguous()


def _build_kmap(coords_in, coords_out, offsets):
    refs = sphash(coords_in)
    res = sphashquery(sphash(coords_out, offsets), refs)  # [K, N_out] -> input index or -1
    return res


def _apply_kmap(feats, weight, res, n_out, transposed):
    """Gather - GEMM - scatter-add, one kernel offset at a time (index order, deterministic).
    res[k, j] = input row feeding output j through offset k (non-transposed)."""
    K = res.shape[0]
    out = torch.zeros(n_out, weight.shape[-1], dtype=feats.dtype)
    for k in range(K):
        rk = res[k]
        j = torch.nonzero(rk >= 0).squeeze(1)
        if j.numel() == 0:
            continue
        i = rk[j]
        if not transposed:
            out.index_add_(0, j, feats[i] @ weight[k])
        else:  # roles swapped: features live on the (coarse) output side of the stored map
            out.index_add_(0, i, feats[j] @ weight[k])
    return out


def conv3d(input, weight, bias=None, kernel_size=3, stride=1, dilation=1, transposed=False):
    feats, coords = input.feats, input.coords
    kernel_size, stride, dilation = _t(kernel_size), _t(stride), _t(dilation)
    if kernel_size == (1, 1, 1) and stride == (1, 1, 1) and dilation == (1, 1, 1):
        out_f = feats.matmul(weight)
        if bias is not None:
            out_f = out_f + bias
        output = SparseTensor(out_f, coords, input.stride)
    elif not transposed:
        key = (input.stride, kernel_size, stride, dilation)
        kmap = input.kmaps.get(key)
        if kmap is None:
            offsets = get_kernel_offsets(kernel_size, stride=input.stride)
            out_coords = coords
            if any(s > 1 for s in stride):
                out_coords = spdownsample(coords, stride, kernel_size, input.stride)
            res = _build_kmap(coords, out_coords, offsets)
            kmap = (res, out_coords, (feats.shape[0], out_coords.shape[0]))
            input.kmaps[key] = kmap
        res, out_coords, sizes = kmap
        out_f = _apply_kmap(feats, weight, res, sizes[1], False)
        if bias is not None:
            out_f = out_f + bias
        output = SparseTensor(out_f, out_coords, tuple(input.stride[k] * stride[k] for k in range(3)))
    else:
        ts = tuple(input.stride[k] // stride[k] for k in range(3))
        res, _, sizes = input.kmaps[(ts, kernel_size, stride, dilation)]
        out_f = _apply_kmap(feats, weight, res, sizes[0], True)
        if bias is not None:
            out_f = out_f + bias
        output = SparseTensor(out_f, input.cmaps[ts], ts)
    output.cmaps = input.cmaps
    output.cmaps.setdefault(output.stride, output.coords)
    output.kmaps = input.kmaps
    return output
