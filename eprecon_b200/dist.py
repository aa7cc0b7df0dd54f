"""Multi-GPU: fragment-per-rank data parallelism + the one exchange step of the path — gathering every rank's final
sparse TSDF into the HOLDER of the global volume (BASELINE.json configs[3]).

The reference has no collective at inference (each rank's GRUFusion state is process-local, models/gru_fusion.py:31-38);
the merge rule is derived from GRUFusion(direct_substitute=True) (models/gru_fusion.py:93-94,198-204): a fragment
REPLACES every global voxel inside its bounding volume.  Fragments are applied in (rank, slot) order, so the result is
deterministic.

Exchange, per step:
  1. all_gather of a tiny header (per fragment: row count + box) — the only thing every rank receives;
  2. every non-holder rank isends its rows ((x, y, z, tsdf bits) int32 [n,4], UNPADDED) to the holder, which posts the
     matching irecvs into one contiguous buffer (grouped NCCL send/recv over NVLink; gloo on CPU);
  3. the holder runs ONE kernel over all rows (`ep_merge_substitute_flags`: a row survives iff no later fragment's box
     contains it) + stable compaction + one int4 row gather.
Round 1 all-gathered a padded payload to every rank and ran a per-rank Python mask/cat loop on all of them (O(N^2) per
step, host-synchronising): that was the limiter of the 1->8 curve.
"""
import torch
import torch.distributed as dist


def pack_rows(coords_global, tsdf):
    """coords_global int [n,3] (global 4 cm voxel ids = local + relative origin), tsdf fp32 [n] or [n,1] -> int32 [n,4]."""
    n = coords_global.shape[0]
    rows = torch.empty((n, 4), dtype=torch.int32, device=tsdf.device)
    rows[:, :3] = coords_global
    rows[:, 3] = tsdf.reshape(-1).float().view(torch.int32)
    return rows


def unpack_rows(rows):
    return rows[:, :3], rows[:, 3].view(torch.float32)


def gather_to_holder(rows_list, boxes, dst=0, group=None):
    """rows_list: this rank's fragments, each int32 [n_s,4] (pack_rows); boxes: per fragment (lo [3], hi [3]) global voxel
    bounds (hi exclusive).  Every rank must pass the same number of fragments.  Returns on the holder a dict
    {rows int32 [N,4] (rank-major, slot-minor), frag_start int32 [F+1] (device), boxes int32 [F,6] (device), counts
    [[n per slot] per rank]}, on the other ranks None."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = rows_list[0].device
    S = len(rows_list)
    header = torch.tensor([[r.shape[0], *lo, *hi] for r, (lo, hi) in zip(rows_list, boxes)], dtype=torch.int32)   # host: sizes are known
    header = header.to(dev, non_blocking=True)
    if world > 1:
        headers = [torch.empty_like(header) for _ in range(world)]
        dist.all_gather(headers, header, group=group)
        headers = torch.stack(headers)                      # [world, S, 7]
    else:
        headers = header.unsqueeze(0)
    mine = rows_list[0] if S == 1 else torch.cat(rows_list)
    if rank != dst:
        if mine.shape[0] > 0:
            for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, mine, dst, group)]):
                req.wait()
        return None
    counts = headers[:, :, 0].tolist()                      # the holder's one read-back: receive sizes
    per_rank = [sum(c) for c in counts]
    total = sum(per_rank)
    buf = torch.empty((total, 4), dtype=torch.int32, device=dev)
    ops_, off = [], 0
    for r in range(world):
        if r == dst:
            buf[off:off + per_rank[r]].copy_(mine)
        elif per_rank[r] > 0:
            ops_.append(dist.P2POp(dist.irecv, buf[off:off + per_rank[r]], r, group))
        off += per_rank[r]
    if ops_:
        for req in dist.batch_isend_irecv(ops_):
            req.wait()
    flat = headers.view(world * S, 7)
    frag_start = torch.zeros(world * S + 1, dtype=torch.int32, device=dev)
    frag_start[1:] = torch.cumsum(flat[:, 0], 0)
    return {"rows": buf, "frag_start": frag_start, "boxes": flat[:, 1:].contiguous(), "counts": counts}


def merge_substitute(rows, frag_start, boxes):
    """In-order substitute-inside-bounding-volume merge of gathered fragments on the GPU (one flags kernel + stable
    compaction + one int4 gather).  rows int32 [N,4]; frag_start int32 [F+1]; boxes int32 [F,6].  -> merged rows [M,4]."""
    from . import _lib, ops
    if not rows.is_cuda:
        raise _lib.EpreconError("merge_substitute runs on the GPU (eprecon_b200 has no CPU path)")
    n = rows.shape[0]
    if n == 0:
        return rows
    flags = torch.empty(n, dtype=torch.uint8, device=rows.device)
    _lib.check(_lib.lib().ep_merge_substitute_flags(rows.data_ptr(), n, frag_start.data_ptr(), boxes.data_ptr(),
                                                    boxes.shape[0], flags.data_ptr(), ops.stream_ptr()),
               "ep_merge_substitute_flags")
    index, _ = ops.compact_flags(flags)
    return ops.gather_coords(rows, index)
