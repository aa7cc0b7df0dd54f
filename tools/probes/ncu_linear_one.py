"""Two dense-linear shapes of the bench fragment through both kernels (linear_rows_kernel, then spconv_kernel via an identity
neighbour table) for an `ncu --set full` capture:

    ncu --set full --clock-control none --import-source on -k regex:"linear_rows|spconv_kernel" -c 4 -o gpurun_out/r02_linear_prof \
        python tools/probes/ncu_linear_one.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from eprecon_b200 import _lib, ops  # noqa: E402

L = _lib.lib()
for m, cin, cout in ((209800, 48, 24), (105000, 48, 192)):
    g = torch.Generator().manual_seed(cin)
    x = torch.randn(m, ops.ceil4(cin), generator=g).cuda()
    W = (torch.randn(1, cin, ops.ceil4(cout), generator=g) / cin ** 0.5).cuda()
    out = torch.empty(m, ops.ceil4(cout), device="cuda")
    part = torch.empty(L.ep_spconv_num_row_tiles(m), 2, cout, device="cuda")
    nbr = torch.arange(m, dtype=torch.int32, device="cuda").view(m, 1).contiguous()
    for nb in (0, nbr.data_ptr()):
        torch.cuda.synchronize()
        _lib.check(L.ep_spconv_fwd(x.data_ptr(), x.stride(0), cin, nb, 1, W.data_ptr(), W.shape[2], cout, 0, out.data_ptr(),
                                   out.stride(0), m, part.data_ptr(), ops.stream_ptr()), "fwd")
        torch.cuda.synchronize()
