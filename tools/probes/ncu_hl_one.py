"""Two launches of the dominant kernel at the bench fragment's level-2 sizes (74->8 stem, 48->24 ConvGRU conv; ~197 k rows,
4.1 M neighbour pairs, Morton row order) for an `ncu --set full` capture:

    ncu --set full --clock-control none --import-source on -k regex:spconv_hl_cp -c 2 -o gpurun_out/r02_hl_prof \
        python tools/probes/ncu_hl_one.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from eprecon_b200 import _lib, ops  # noqa: E402
from eprecon_b200.sparse import VoxelSet  # noqa: E402
from probe_hl import slab_set  # noqa: E402

L = _lib.lib()
coords = slab_set(182).cuda()
nbr = VoxelSet(coords, 1).kmap_k3()
m = coords.shape[0]
for cin, cout in ((74, 8), (48, 24)):
    g = torch.Generator().manual_seed(cin)
    x = torch.zeros(m, ops.ceil4(cin))
    x[:, :cin] = torch.randn(m, cin, generator=g)
    W = torch.zeros(27, cin, ops.ceil4(cout))
    W[:, :, :cout] = torch.randn(27, cin, cout, generator=g) / (27 * cin) ** 0.5
    xc, Wc = x.cuda(), W.cuda()
    w_hl, npad = ops._hl_weights(Wc, cout)
    x_hl = ops.hl_split(xc, cin)
    o = torch.empty((m, ops.ceil4(cout)), dtype=torch.float32, device="cuda")
    part = torch.empty((L.ep_spconv_num_row_tiles(m), 2, cout), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    _lib.check(L.ep_spconv_hl_fwd(x_hl.data_ptr(), m, cin, nbr.data_ptr(), 27, w_hl.data_ptr(), npad, cout, 0, o.data_ptr(),
                                  o.stride(0), m, part.data_ptr(), 0, 0, 0, ops.stream_ptr()), "hl")
    torch.cuda.synchronize()
    pairs = int((nbr >= 0).sum())
    print(f"cin {cin} cout {cout} rows {m} pairs {pairs} gathered bytes {pairs * ((cin + 31) // 32) * 128}")
