"""world_size-2 gloo tests of the global-TSDF gather-to-holder (run on CPU) + the merge rule against the oracle.

The merge kernel itself is CUDA (tests/test_dist_gpu.py); here the gathered rows are merged by `merge_reference`, a
plain-torch restatement of the same rule, and both are pinned to oracle/restate.gru_fusion(direct_substitute=True) --
the reference's GRUFusion(direct_substitute=True) semantics (models/gru_fusion.py:93-94,198-204) -- on their ACTIVE
voxels (|tsdf| < 1): the reference additionally keeps dead rows (tsdf = 1) where only the old scene had a voxel."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def merge_reference(rows, frag_start, boxes):
    """Per-fragment mask + cat, the literal in-order rule (what round 1 ran on every rank)."""
    g = torch.zeros((0, 4), dtype=torch.int32)
    fs = frag_start.tolist()
    for f in range(len(fs) - 1):
        cur = rows[fs[f]:fs[f + 1]]
        if cur.shape[0] == 0:
            continue
        lo, hi = boxes[f, :3], boxes[f, 3:]
        inside = ((g[:, :3] >= lo) & (g[:, :3] < hi)).all(-1)
        g = torch.cat([g[~inside], cur])
    return g


def scene_fragments(n):
    """n overlapping fragments of one synthetic scene: (rows int32 [k,4] global, box lo, box hi, oracle inputs...)."""
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_golden_panoptic_fusion import N_VOX, fusion_inputs
    out = []
    for frag in range(n):
        inputs, coords, tsdf, _ = fusion_inputs(frag)
        rel = ((inputs["vol_origin_partial"][0] - inputs["vol_origin"][0]) / 0.04).long()
        out.append({"inputs": inputs, "coords": coords, "tsdf": tsdf, "rel": rel,
                    "lo": rel.tolist(), "hi": (rel + torch.tensor(N_VOX)).tolist()})
    return out, N_VOX


def oracle_active_set(frags, n_vox):
    """restate.gru_fusion(direct_substitute=True) over the fragments in order -> sorted active (x,y,z,tsdf bits) rows."""
    from oracle import restate
    from eprecon_b200 import synth
    cfg = synth.make_cfg(n_vox=n_vox)
    state = restate.FusionState()
    for fr in frags:
        ins = {k: v for k, v in fr["inputs"].items() if k not in ("occ_list", "tsdf_list")}
        restate.gru_fusion(state, None, cfg, fr["coords"], fr["tsdf"], ins, 2, [96, 48, 24], direct_substitute=True)
    C, F = state.C[2], state.F[2][:, 0]
    act = F.abs() < 1
    return canonical(C[act].int(), F[act])


def canonical(c, f):
    key = (c[:, 0].long() * 100000 + c[:, 1].long()) * 100000 + c[:, 2].long()
    o = torch.argsort(key)
    return c[o].int(), f[o].float()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from eprecon_b200.dist import gather_to_holder, pack_rows, unpack_rows
    frags, n_vox = scene_fragments(4)
    mine = frags[2 * rank:2 * rank + 2]                 # 2 fragments ("streams") per rank, rank-major order
    rows = [pack_rows(fr["coords"][:, 1:].int() + fr["rel"].int(), fr["tsdf"]) for fr in mine]
    got = gather_to_holder(rows, [(fr["lo"], fr["hi"]) for fr in mine], dst=0)
    if rank == 0:
        c, f = unpack_rows(got["rows"])
        q.put((rank, got["counts"], got["frag_start"].tolist(), got["boxes"].tolist(), got["rows"]))
    else:
        q.put((rank, got is None, None, None, None))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_to_holder_world2_matches_oracle_merge():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted((q.get(timeout=180) for _ in ps), key=lambda t: t[0])
    for p in ps:
        p.join(60)
    assert res[1][1] is True                                   # only the holder receives anything
    _, counts, frag_start, boxes, rows = res[0]
    frags, n_vox = scene_fragments(4)
    want_counts = [[len(frags[0]["coords"]), len(frags[1]["coords"])], [len(frags[2]["coords"]), len(frags[3]["coords"])]]
    assert counts == want_counts
    assert frag_start[-1] == rows.shape[0] == sum(sum(c) for c in counts)
    # payload round-trips bit-exactly, rank-major / slot-minor
    from eprecon_b200.dist import pack_rows
    want_rows = torch.cat([pack_rows(fr["coords"][:, 1:].int() + fr["rel"].int(), fr["tsdf"]) for fr in frags])
    assert torch.equal(rows, want_rows)
    assert boxes == [fr["lo"] + fr["hi"] for fr in frags]
    merged = merge_reference(rows, torch.tensor(frag_start), torch.tensor(boxes, dtype=torch.int32))
    mc, mf = merged[:, :3], merged[:, 3].view(torch.float32)
    act = mf.abs() < 1
    gc, gf = canonical(mc[act], mf[act])
    oc, of = oracle_active_set(frags, n_vox)
    assert torch.equal(gc, oc) and torch.equal(gf, of)         # active scene voxels and values: bit-exact vs the oracle


def test_merge_rule_small():
    rows = torch.tensor([[0, 0, 0, 1], [50, 1, 1, 2], [50, 1, 1, 7], [60, 2, 2, 8]], dtype=torch.int32)
    fs = torch.tensor([0, 2, 4])
    boxes = torch.tensor([[0, 0, 0, 96, 96, 96], [40, 0, 0, 136, 96, 96]], dtype=torch.int32)
    m = merge_reference(rows, fs, boxes)
    assert m.tolist() == [[0, 0, 0, 1], [50, 1, 1, 7], [60, 2, 2, 8]]
