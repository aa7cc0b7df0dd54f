"""tcgen05 tensor-core sparse convs (csrc/spconv_hl.cu: half-pair operands, cp.async gather; csrc/spconv_tc.cu: 3xTF32, register
producers) vs the fp32 CUDA-core kernel and fp64."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def _case(m_in, m_out, cin, cout, K, seed, density=0.5):
    g = torch.Generator().manual_seed(seed)
    from eprecon_b200 import ops
    x = torch.zeros(m_in, ops.ceil4(cin))
    x[:, :cin] = torch.randn(m_in, cin, generator=g)
    if K == 1:
        nbr = None
        m_out = m_in
    else:
        nbr = torch.randint(0, m_in, (m_out, K), generator=g, dtype=torch.int32)
        nbr[torch.rand(m_out, K, generator=g) > density] = -1
    W = torch.zeros(K, cin, ops.ceil4(cout))
    W[:, :, :cout] = torch.randn(K, cin, cout, generator=g) / (K * cin) ** 0.5
    bias = torch.randn(cout, generator=g)
    # fp64 ground truth
    want = torch.zeros(m_out, cout, dtype=torch.float64)
    xd, Wd = x[:, :cin].double(), W[:, :, :cout].double()
    for k in range(K):
        if nbr is None:
            want += xd @ Wd[k]
        else:
            ok = nbr[:, k] >= 0
            want[ok] += xd[nbr[ok, k].long()] @ Wd[k]
    want += bias.double()
    return x, nbr, W, bias, want.float(), m_out


@pytest.mark.parametrize("m_in,m_out,cin,cout,K", [
    (300, 257, 16, 16, 27), (1000, 900, 80, 32, 27), (5000, 4100, 138, 16, 27), (700, 650, 32, 1, 27),
    (2000, 2000, 96, 384, 1), (900, 300, 64, 64, 8), (400, 400, 160, 96, 1), (3000, 2500, 24, 24, 27),
    (129, 129, 8, 8, 27), (640, 640, 40, 40, 27)])
def test_tc_matches_ffma_and_fp64(cuda_lib, m_in, m_out, cin, cout, K):
    from eprecon_b200 import ops
    x, nbr, W, bias, want, m_out = _case(m_in, m_out, cin, cout, K, seed=cin * 7 + cout)
    xc, Wc, bc = x.cuda(), W.cuda(), bias.cuda()
    nc = nbr.cuda() if nbr is not None else None
    outs, parts = {}, {}
    old = ops.SPCONV_IMPL
    try:
        for impl in ("ffma", "tf32x3", "tf32", "hl"):
            ops.SPCONV_IMPL = impl
            y, part = ops.spconv(xc, cin, nc, Wc, cout, bias=bc, m_out=m_out, want_stats=True)
            torch.cuda.synchronize()
            outs[impl], parts[impl] = y[:, :cout].cpu(), part.cpu()
    finally:
        ops.SPCONV_IMPL = old
    assert rel(outs["ffma"], want) < 2e-6
    assert rel(outs["tf32x3"], want) < 3e-5, rel(outs["tf32x3"], want)      # fp32-grade (2^-21 split residue)
    assert rel(outs["tf32"], want) < 3e-3, rel(outs["tf32"], want)          # single-pass tf32
    if K > 1:
        assert rel(outs["hl"], want) < 3e-5, rel(outs["hl"], want)          # half-pair operands: 22 significant bits, like 3xTF32
    # fused BatchNorm statistics agree with the column sums of the output
    for impl in ("ffma", "tf32x3", "hl"):
        s = parts[impl][:, 0].sum(0)
        q = parts[impl][:, 1].sum(0)
        assert torch.allclose(s, outs[impl].sum(0), rtol=1e-4, atol=1e-3), impl
        assert torch.allclose(q, (outs[impl] ** 2).sum(0), rtol=1e-4, atol=1e-3), impl
