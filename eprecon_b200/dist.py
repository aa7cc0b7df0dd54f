"""Multi-GPU: fragment-per-rank data parallelism + the one exchange step of the path — gathering every rank's final
sparse TSDF into the holder of the global volume (BASELINE.json config 4).

The reference has no collective at inference (each rank's GRUFusion state is process-local, models/gru_fusion.py:31-38);
the merge rule is derived from GRUFusion(direct_substitute=True) (models/gru_fusion.py:93-94,198-204): a fragment
REPLACES every global voxel inside its bounding volume.  Ranks are merged in rank order, so the result is deterministic.

Works on NCCL (GPU) and gloo (CPU tests, world_size 2).
"""
import torch
import torch.distributed as dist


def gather_fragments(coords_global, tsdf, dst=0, group=None):
    """coords_global int32 [n,3] (global 4 cm voxel ids = local + relative origin), tsdf fp32 [n].
    all_gather of the counts (world x 8 B) then one padded all_gather of the payload (<= ~1.9 MB/rank).
    Returns on every rank a list of (coords, tsdf) per source rank (rank order)."""
    world = dist.get_world_size(group)
    dev = tsdf.device
    n = torch.tensor([coords_global.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    payload = torch.zeros((cap, 4), dtype=torch.int32, device=dev)
    payload[:n.item(), :3] = coords_global.to(torch.int32)
    payload[:n.item(), 3] = tsdf.float().view(torch.int32) if tsdf.numel() else 0
    bufs = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(bufs, payload, group=group)
    return [(b[:c, :3], b[:c, 3].view(torch.float32)) for b, c in zip(bufs, counts)]


def merge_substitute(fragments, boxes):
    """Merge per-rank sparse TSDFs in rank order with the reference's substitute-inside-bounding-volume rule.
    fragments: list of (coords int32 [n,3], tsdf [n]); boxes: list of (lo [3], hi [3]) global voxel bounds (hi exclusive)."""
    dev = fragments[0][0].device
    gC = torch.zeros((0, 3), dtype=torch.int32, device=dev)
    gF = torch.zeros((0,), dtype=torch.float32, device=dev)
    for (c, f), (lo, hi) in zip(fragments, boxes):
        lo_t = torch.as_tensor(lo, dtype=torch.int32, device=dev)
        hi_t = torch.as_tensor(hi, dtype=torch.int32, device=dev)
        inside = ((gC >= lo_t) & (gC < hi_t)).all(-1)
        gC = torch.cat([gC[~inside], c])
        gF = torch.cat([gF[~inside], f])
    return gC, gF
