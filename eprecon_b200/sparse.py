"""Device-side sparse coordinate structures: what torchsparse keeps in SparseTensor.cmaps / .kmaps and
PointTensor.idx_query / .weights (reference: ops/torchsparse_utils.py:15-105), rebuilt B200-first:

  * VoxelSet  — int32 coords [M,4]=(x,y,z,b) at a tensor stride + ONE open-addressing hash table; every
                kernel map on the set (k3 neighbours, k2s2 down / transposed up) is an output-major
                table built once and cached, so each conv is an atomics-free gather-GEMM.
  * PointCloud — float points (already divided by the voxel resolution), their voxelisation CSR
                (deterministic segmented mean instead of float atomics) and cached trilinear taps.

All arrays live on the GPU; the only host round-trips are the data-dependent sizes (voxel counts).
"""
import torch

from . import _lib, ops
from .ops import stream_ptr, _ptr

_L = _lib.lib


class VoxelSet:
    def __init__(self, coords, stride, keys=None):
        self.coords = coords            # int32 [M,4] (x,y,z,b)
        self.stride = stride
        self.m = coords.shape[0]
        if keys is None:
            keys = ops.coord_keys(coords, batch_first=False)
        self.table = ops.HashTable(keys)
        self._k3 = None
        self._down = None

    def kmap_k3(self):
        """nbr [M,27]: row of coord + off*stride, offsets z-outer/x-inner (torchsparse odd-kernel weight order)."""
        if self._k3 is None:
            offs = ops.kernel_offsets("k3", self.stride, self.coords.device)
            self._k3 = ops.kmap_build(self.coords, False, offs, self.table)
        return self._k3

    def downsample(self):
        """k2s2 strided map (spdownsample, k == s): returns (coarse VoxelSet, nbr_down [Mc,8], nbr_up [M,8])
        where nbr_up has at most one valid entry per fine row (the transposed conv as a gather)."""
        if self._down is None:
            dev = self.coords.device
            L = _L()
            step = 2 * self.stride
            keys = torch.empty(self.m, dtype=torch.int64, device=dev)
            _lib.check(L.ep_down_keys(self.coords.data_ptr(), self.m, step, keys.data_ptr(), stream_ptr()), "ep_down_keys")
            seg = ops.sort_segments(keys, 64, want_seg_of_item=False)
            mc = seg["S"]
            cc = torch.empty((mc, 4), dtype=torch.int32, device=dev)
            _lib.check(L.ep_down_unpack(seg["keys_sorted"].data_ptr(), seg["seg_start"].data_ptr(), mc, cc.data_ptr(),
                                        stream_ptr()), "ep_down_unpack")
            coarse = VoxelSet(cc, step)
            offs = ops.kernel_offsets("k2", self.stride, dev)
            nbr_down = ops.kmap_build(cc, False, offs, self.table)           # [Mc,8] -> fine rows
            # one-hot inverse [M,8] (at most one valid entry per fine row): the same gather-GEMM kernel runs the transposed conv
            nbr_up = torch.full((self.m, 8), -1, dtype=torch.int32, device=dev)
            _lib.check(L.ep_kmap_inverse(nbr_down.data_ptr(), mc, 8, nbr_up.data_ptr(), stream_ptr()), "ep_kmap_inverse")
            self._down = (coarse, nbr_down, nbr_up)
        return self._down


class PointCloud:
    """Points + their stride-1 voxelisation (initial_voxelize, ops/torchsparse_utils.py:15-35)."""

    def __init__(self, pts, vres, order="spatial"):
        """pts float32 [N,4]=(x,y,z,b) in metres (aligned-camera frame); vres = voxel resolution.
        order: "spatial" = voxel rows in (b,x,y,z) raster order (neighbour gathers stay cache-local; every consumer is
        order-independent), "hash" = the reference's ascending-sphash order (torch.unique of the hashes), needed only
        where the reference's ConvGRU.convr quirk makes the row order observable."""
        dev = pts.device
        L = _L()
        n = pts.shape[0]
        self.n = n
        self.scaled = torch.empty_like(pts)
        keys = torch.empty(n, dtype=torch.int64, device=dev)
        spatial = order == "spatial"
        _lib.check(L.ep_point_keys(pts.data_ptr(), n, float(vres), int(spatial), self.scaled.data_ptr(), keys.data_ptr(),
                                   stream_ptr()), "ep_point_keys")
        seg = ops.sort_segments(keys, 64 if spatial else 60)
        self.csr1 = (seg["perm"], seg["seg_start"], seg["seg_end"])
        self.idx_query = seg["seg_of_item"]
        m = seg["S"]
        vox = torch.empty((m, 4), dtype=torch.int32, device=dev)
        _lib.check(L.ep_segment_coords(self.scaled.data_ptr(), seg["perm"].data_ptr(), seg["seg_start"].data_ptr(), m,
                                       vox.data_ptr(), stream_ptr()), "ep_segment_coords")
        self.vox = VoxelSet(vox, 1, keys=None if spatial else seg["keys_sorted"][seg["seg_start"].long()])
        self.order = order
        self._hash_rank = None
        self._taps = {}
        self._csr = {1: self.csr1}

    def voxelize(self, feat, c, csr=None):
        """Mean-pool point rows into voxels of `csr` (default: the stride-1 voxelisation)."""
        perm, s0, s1 = csr if csr is not None else self.csr1
        m = s0.shape[0]
        out = torch.empty((m, ops.ceil4(c)), dtype=torch.float32, device=feat.device)
        _lib.check(_L().ep_segment_mean(feat.data_ptr(), feat.stride(0), c, perm.data_ptr(), s0.data_ptr(), s1.data_ptr(),
                                        m, out.data_ptr(), out.stride(0), stream_ptr()), "ep_segment_mean")
        return out

    def taps(self, vset):
        """(idx [N,8], w [N,8]) trilinear taps of every point into `vset` (voxel_to_point, cached per stride)."""
        if vset.stride not in self._taps:
            idx = torch.empty((self.n, 8), dtype=torch.int32, device=self.scaled.device)
            w = torch.empty((self.n, 8), dtype=torch.float32, device=self.scaled.device)
            _lib.check(_L().ep_devox_prepare(self.scaled.data_ptr(), self.n, vset.stride, vset.table.keys.data_ptr(),
                                             vset.table.vals.data_ptr(), vset.table.cap, idx.data_ptr(), w.data_ptr(),
                                             stream_ptr()), "ep_devox_prepare")
            self._taps[vset.stride] = (idx, w)
        return self._taps[vset.stride]

    def taps_in_hash_order(self):
        """Stride-1 taps with the voxel ids translated to the reference's ascending-hash row order (what the reference's
        cached idx_query holds when ConvGRU.convr reuses convz's taps on a different voxel list)."""
        idx, w = self.taps(self.vox)
        if self.order == "hash":
            return idx, w
        if self._hash_rank is None:
            seg = ops.sort_segments(ops.coord_keys(self.vox.coords, batch_first=False), 60)
            rank = seg["seg_of_item"]                       # hash rank of every (spatially ordered) voxel row
            flat = idx.view(-1).long()
            tr = torch.where(flat >= 0, rank[flat.clamp_min(0)], torch.full_like(flat, -1, dtype=torch.int32))
            self._hash_rank = tr.view(-1, 8).contiguous()
        return self._hash_rank, w

    def csr_for(self, vset):
        """CSR of points by voxel of `vset` (point_to_voxel, ops/torchsparse_utils.py:40-63); points whose cell is
        not an active voxel are dropped, as in the reference (idx_query == -1)."""
        if vset.stride not in self._csr:
            dev = self.scaled.device
            idx = torch.empty(self.n, dtype=torch.int32, device=dev)
            keys = torch.empty(self.n, dtype=torch.int64, device=dev)
            _lib.check(_L().ep_point_query(self.scaled.data_ptr(), self.n, vset.stride, vset.table.keys.data_ptr(),
                                           vset.table.vals.data_ptr(), vset.table.cap, idx.data_ptr(), keys.data_ptr(),
                                           vset.m, stream_ptr()), "ep_point_query")
            seg = ops.sort_segments(keys, max(1, int(vset.m).bit_length()), sentinel=vset.m, want_seg_of_item=False)
            # segments only exist for voxels that received >= 1 point; expand to all voxels of vset
            s0 = torch.zeros(vset.m, dtype=torch.int32, device=dev)
            s1 = torch.zeros(vset.m, dtype=torch.int32, device=dev)
            ids = seg["keys_sorted"][seg["seg_start"].long()].long()
            s0[ids] = seg["seg_start"]
            s1[ids] = seg["seg_end"]
            self._csr[vset.stride] = (seg["perm"], s0, s1)
        return self._csr[vset.stride]


def devoxelize(vfeat, c, idx, w, add=None, out=None):
    n = idx.shape[0]
    if out is None:
        out = torch.empty((n, ops.ceil4(c)), dtype=torch.float32, device=vfeat.device)
    _lib.check(_L().ep_devoxelize(vfeat.data_ptr(), vfeat.stride(0), c, idx.data_ptr(), w.data_ptr(), n, _ptr(add),
                                  add.stride(0) if add is not None else 0, out.data_ptr(), out.stride(0), stream_ptr()),
               "ep_devoxelize")
    return out
