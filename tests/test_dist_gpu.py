"""The holder-side merge kernel (csrc/union_gather.cu::ep_merge_substitute_flags) against the literal per-fragment rule
and the oracle's GRUFusion(direct_substitute=True); plus the 2-rank NCCL gather + merge when the box has >= 2 GPUs."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from test_dist_cpu import canonical, merge_reference, oracle_active_set, scene_fragments  # noqa: E402


def test_merge_kernel_matches_reference_rule_random(cuda_lib):
    from eprecon_b200.dist import merge_substitute
    g = torch.Generator().manual_seed(3)
    rows, fs, boxes = [], [0], []
    for f in range(37):
        lo = torch.randint(0, 200, (3,), generator=g)
        n = 0 if f in (5, 20) else int(torch.randint(1, 4000, (1,), generator=g))    # two EMPTY fragments
        c = torch.randint(0, 96, (n, 3), generator=g) + lo
        t = torch.randint(-2 ** 31, 2 ** 31 - 1, (n, 1), generator=g, dtype=torch.int64)
        rows.append(torch.cat([c, t], 1).int())
        fs.append(fs[-1] + n)
        boxes.append(lo.tolist() + (lo + 96).tolist())
    rows = torch.cat(rows)
    fs_t, bx_t = torch.tensor(fs, dtype=torch.int32), torch.tensor(boxes, dtype=torch.int32)
    want = merge_reference(rows, fs_t, bx_t)
    got = merge_substitute(rows.cuda(), fs_t.cuda(), bx_t.cuda()).cpu()
    assert torch.equal(got, want)                          # same rows, same (stable) order


def test_merge_kernel_matches_oracle_direct_substitute(cuda_lib):
    from eprecon_b200.dist import merge_substitute, pack_rows
    frags, n_vox = scene_fragments(4)
    rows = torch.cat([pack_rows(fr["coords"][:, 1:].int() + fr["rel"].int(), fr["tsdf"]) for fr in frags])
    fs = torch.tensor([0] + list(torch.cumsum(torch.tensor([len(fr["coords"]) for fr in frags]), 0)), dtype=torch.int32)
    bx = torch.tensor([fr["lo"] + fr["hi"] for fr in frags], dtype=torch.int32)
    m = merge_substitute(rows.cuda(), fs.cuda(), bx.cuda()).cpu()
    mf = m[:, 3].view(torch.float32)
    act = mf.abs() < 1
    gc, gf = canonical(m[act, :3], mf[act])
    oc, of = oracle_active_set(frags, n_vox)
    assert torch.equal(gc, oc) and torch.equal(gf, of)


def _nccl_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from eprecon_b200.dist import gather_to_holder, merge_substitute, pack_rows
    frags, n_vox = scene_fragments(4)
    mine = frags[2 * rank:2 * rank + 2]
    rows = [pack_rows((fr["coords"][:, 1:].int() + fr["rel"].int()).cuda(), fr["tsdf"].cuda()) for fr in mine]
    got = gather_to_holder(rows, [(fr["lo"], fr["hi"]) for fr in mine], dst=0)
    if rank == 0:
        m = merge_substitute(got["rows"], got["frag_start"], got["boxes"]).cpu()
        q.put(m)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_nccl_gather_and_merge_matches_oracle(cuda_lib):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    m = q.get(timeout=300)
    for p in ps:
        p.join(60)
    frags, n_vox = scene_fragments(4)
    mf = m[:, 3].view(torch.float32)
    act = mf.abs() < 1
    gc, gf = canonical(m[act, :3], mf[act])
    oc, of = oracle_active_set(frags, n_vox)
    assert torch.equal(gc, oc) and torch.equal(gf, of)
