"""Dense per-row linear (ep_spconv_fwd with K == 1 and no neighbour table -> linear_mma_kernel, csrc/linear_mma.cu: 3xTF32 on
the tensor cores, rows straight from global memory into MMA fragments): the nn.Linear layers of the path (reference:
models/modules.py:127-136,187,279-284).  Checked against
  * an fp64 matmul of the same operands: |err| <= 2e-6 * scale * sqrt(cin) -- fp32-class accuracy, the tolerance of the FFMA kernel,
  * the fp32 FFMA gather-GEMM tile kernel fed an identity neighbour table,
  * the per-64-row-tile column sums / sums of squares the BatchNorm finalisation consumes."""
import pytest
import torch

from eprecon_b200 import _lib, ops

pytestmark = [pytest.mark.gpu]

SHAPES = [(1, 1), (3, 4), (8, 32), (24, 96), (32, 24), (48, 24), (74, 8), (96, 24), (96, 48), (192, 48), (24, 21), (138, 130), (17, 40)]


def run_linear(L, x, cin, W, cout, bias, m, nbr=None, stats=True, out_ld=None, out_col=0):
    c4 = ops.ceil4(cout)
    out_ld = out_ld or c4
    out = torch.full((m, out_ld), -7.0, dtype=torch.float32, device="cuda")
    part = torch.full((L.ep_spconv_num_row_tiles(m), 2, cout), float("nan"), dtype=torch.float32, device="cuda") if stats else None
    _lib.check(L.ep_spconv_fwd(x.data_ptr(), x.stride(0), cin, nbr.data_ptr() if nbr is not None else 0, 1, W.data_ptr(), W.shape[2],
                               cout, bias.data_ptr() if bias is not None else 0, out.data_ptr() + 4 * out_col, out.stride(0), m,
                               part.data_ptr() if stats else 0, ops.stream_ptr()), "ep_spconv_fwd")
    torch.cuda.synchronize()
    return out, part


@pytest.mark.parametrize("cin,cout", SHAPES)
@pytest.mark.parametrize("m", [1, 63, 64, 257, 5000])
def test_linear_rows_matches_matmul_and_tile_kernel(cuda_lib, cin, cout, m):
    L = cuda_lib
    g = torch.Generator().manual_seed(cin * 1000 + cout + m)
    ld = ops.ceil4(cin) + 4                                    # wider than the used channels: a column slice of a concat buffer
    x = torch.zeros(m, ld)
    x[:, :cin] = torch.randn(m, cin, generator=g)
    x[:, ops.ceil4(cin):] = float("nan")                       # columns past ceil4(cin) are never read
    W = torch.zeros(1, cin, ops.ceil4(cout))
    W[0, :, :cout] = torch.randn(cin, cout, generator=g) / cin ** 0.5
    bias = torch.randn(cout, generator=g)
    xc, Wc, bc = x.cuda(), W.cuda(), bias.cuda()
    out, part = run_linear(L, xc, cin, Wc, cout, bc, m)
    want = x[:, :cin].double() @ W[0, :, :cout].double() + bias.double()
    got = out[:, :cout].cpu().double()
    assert (got - want).abs().max() <= 2e-6 * max(1.0, want.abs().max().item()) * max(1, cin) ** 0.5
    if ops.ceil4(cout) != cout:
        assert bool((out[:, cout:] == -7.0).all())             # padding columns are the caller's
    # identity neighbour table -> spconv_kernel (fp32 FFMA gather-GEMM tile)
    nbr = torch.arange(m, dtype=torch.int32, device="cuda").view(m, 1).contiguous()
    ref, ref_part = run_linear(L, xc, cin, Wc, cout, bc, m, nbr=nbr)
    assert (out[:, :cout] - ref[:, :cout]).abs().max().item() <= 2e-6 * max(1.0, want.abs().max().item()) * max(1, cin) ** 0.5
    # per-tile column statistics: exact up to the summation order inside a 64-row tile
    tiles = out[:, :cout].cpu().double().split(64)
    s = torch.stack([t.sum(0) for t in tiles])
    q = torch.stack([(t * t).sum(0) for t in tiles])
    assert torch.allclose(part[:, 0].cpu().double(), s, rtol=1e-5, atol=1e-4)
    assert torch.allclose(part[:, 1].cpu().double(), q, rtol=1e-5, atol=1e-4)
    assert torch.allclose(part.cpu(), ref_part.cpu(), rtol=1e-5, atol=1e-4)


def test_linear_rows_no_bias_no_stats_into_a_column_slice(cuda_lib):
    L = cuda_lib
    g = torch.Generator().manual_seed(9)
    m, cin, cout = 1000, 48, 24
    x = torch.randn(m, cin, generator=g).cuda()
    W = (torch.randn(1, cin, cout, generator=g) / 7).cuda()
    out, _ = run_linear(L, x, cin, W, cout, None, m, stats=False, out_ld=64, out_col=24)   # 96-byte offset: 16-byte aligned or not
    want = (x.double() @ W[0].double()).float()
    assert torch.allclose(out[:, 24:48], want, rtol=1e-5, atol=1e-5)
    assert bool((out[:, :24] == -7.0).all()) and bool((out[:, 48:] == -7.0).all())
    out, _ = run_linear(L, x, cin, W, cout, None, m, stats=False, out_ld=64, out_col=1)    # 4-byte aligned only
    assert torch.allclose(out[:, 1:25], want, rtol=1e-5, atol=1e-5)


def test_linear_rows_large_problem_matches(cuda_lib):
    L = cuda_lib
    g = torch.Generator().manual_seed(11)
    m, cin, cout = 209800, 24, 96
    x = torch.randn(m, cin, generator=g).cuda()
    W = (torch.randn(1, cin, cout, generator=g) / 5).cuda()
    b = torch.randn(cout, generator=g).cuda()
    out, part = run_linear(L, x, cin, W, cout, b, m)
    want = torch.addmm(b.double(), x.double(), W[0].double())
    assert (out.double() - want).abs().max().item() <= 1e-5 * want.abs().max().item()
    assert torch.allclose(part[:, 0].double().sum(0), want.sum(0), rtol=1e-5, atol=1e-2)
