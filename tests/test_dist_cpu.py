"""world_size-2 gloo test of the global-TSDF gather + merge (runs on CPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from eprecon_b200.dist import gather_fragments, merge_substitute
    g = torch.Generator().manual_seed(rank)
    n = 50 + 30 * rank
    lo = torch.tensor([rank * 40, 0, 0])
    coords = (torch.randint(0, 96, (n, 3), generator=g) + lo).int()
    tsdf = torch.rand(n, generator=g)
    frags = gather_fragments(coords, tsdf)
    boxes = [([r * 40, 0, 0], [r * 40 + 96, 96, 96]) for r in range(world)]
    gC, gF = merge_substitute(frags, boxes)
    q.put((rank, [f[0].shape[0] for f in frags], gC.shape[0], float(gF.sum()), frags[rank][1].equal(tsdf)))
    dist.destroy_process_group()


def test_gather_and_merge_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    assert res[0][1] == res[1][1] == [50, 80]                 # every rank sees every fragment
    assert res[0][2] == res[1][2] and abs(res[0][3] - res[1][3]) < 1e-6  # identical merged volume on all ranks
    assert res[0][4] and res[1][4]                             # payload round-trips bit-exactly
    # rank 1's box [40,136) swallows rank-0 voxels with x >= 40: merged count < 130
    assert 80 <= res[0][2] <= 130


def test_merge_substitute_rule():
    from eprecon_b200.dist import merge_substitute
    a = (torch.tensor([[0, 0, 0], [50, 1, 1]], dtype=torch.int32), torch.tensor([0.1, 0.2]))
    b = (torch.tensor([[50, 1, 1], [60, 2, 2]], dtype=torch.int32), torch.tensor([0.7, 0.8]))
    gC, gF = merge_substitute([a, b], [([0, 0, 0], [96, 96, 96]), ([40, 0, 0], [136, 96, 96])])
    assert gC.tolist() == [[0, 0, 0], [50, 1, 1], [60, 2, 2]] and torch.allclose(gF, torch.tensor([0.1, 0.7, 0.8]))
