"""Multi-segment ID GEMM, weight-gradient GEMM and column sums vs fp64 torch on the CPU."""
import pytest
import torch

from graphgym_b200 import ops
from util import FP32_TOL, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(params=['tc', 'simt'], autouse=True)
def gemm_mode(request, monkeypatch):
    """every test runs on the tcgen05 3xTF32 kernel and on the fp32 CUDA-core kernel"""
    monkeypatch.setattr(ops, 'GEMM_MODE', request.param)
    return request.param


@pytest.mark.parametrize('n,k,f', [(1, 1, 1), (37, 12, 16), (300, 1433, 128), (1000, 100, 128),
                                   (513, 128, 256), (129, 7, 130), (64, 1, 128), (2000, 256, 256)])
@pytest.mark.parametrize('b_trans', [False, True])
def test_id_gemm_two_segments(cuda, n, k, f, b_trans):
    g = torch.Generator().manual_seed(n + k + f)
    x = torch.randn(n, k, generator=g)
    w = torch.randn(f, k, generator=g) if b_trans else torch.randn(k, f, generator=g)
    w_id = torch.randn(f, k, generator=g) if b_trans else torch.randn(k, f, generator=g)
    ids = torch.randint(0, n, (max(1, n // 7),), generator=g)  # duplicates on purpose
    cnt = ops.id_count(ids.to(cuda), n)
    assert torch.equal(cnt.cpu(), torch.bincount(ids, minlength=n).float())
    bias = torch.randn(f, generator=g)
    got = ops.id_gemm([(x.to(cuda), w.to(cuda), None), (x.to(cuda), w_id.to(cuda), cnt)], n, f,
                      b_trans=b_trans, bias=bias.to(cuda))
    wd, wid = (w.double().t(), w_id.double().t()) if b_trans else (w.double(), w_id.double())
    want = x.double() @ wd
    want.index_add_(0, ids, x.double()[ids] @ wid)
    want += bias.double()
    assert rel_err(got, want) < FP32_TOL


def test_id_gemm_relu_and_mask(cuda):
    g = torch.Generator().manual_seed(0)
    n, k, f = 333, 64, 96
    x, w = torch.randn(n, k, generator=g), torch.randn(k, f, generator=g)
    m = torch.randn(n, f, generator=g)
    got = ops.id_gemm([(x.to(cuda), w.to(cuda), None)], n, f, act=ops.ACT_RELU)
    assert rel_err(got, (x.double() @ w.double()).relu()) < FP32_TOL
    got = ops.id_gemm([(x.to(cuda), w.to(cuda), None)], n, f, relu_mask=m.to(cuda))
    assert rel_err(got, (x.double() @ w.double()) * (m > 0)) < FP32_TOL


def test_id_gemm_skips_tiles_without_centres(cuda):
    """Centres only in the first rows (the ego-net batch layout): later tiles skip the ID segment."""
    g = torch.Generator().manual_seed(1)
    n, k, f = 1000, 32, 64
    x, w, w_id = torch.randn(n, k, generator=g), torch.randn(k, f, generator=g), torch.randn(k, f, generator=g)
    ids = torch.arange(40)
    cnt = ops.id_count(ids.to(cuda), n)
    got = ops.id_gemm([(x.to(cuda), w.to(cuda), None), (x.to(cuda), w_id.to(cuda), cnt)], n, f)
    want = x.double() @ w.double()
    want[:40] += x.double()[:40] @ w_id.double()
    assert rel_err(got, want) < FP32_TOL


@pytest.mark.parametrize('n,k,f', [(1, 3, 5), (1000, 12, 16), (5000, 1433, 128), (100000, 128, 128),
                                   (777, 100, 256)])
def test_gemm_tn(cuda, n, k, f):
    g = torch.Generator().manual_seed(n)
    a, gr = torch.randn(n, k, generator=g), torch.randn(n, f, generator=g)
    got = ops.gemm_tn(a.to(cuda), gr.to(cuda))
    assert rel_err(got, a.double().t() @ gr.double()) < FP32_TOL
    ids = torch.randint(0, n, (max(1, n // 9),), generator=g)
    got = ops.gemm_tn(a.to(cuda), gr.to(cuda), ids.to(cuda))
    assert rel_err(got, a.double()[ids].t() @ gr.double()[ids]) < FP32_TOL
    assert torch.equal(got, ops.gemm_tn(a.to(cuda), gr.to(cuda), ids.to(cuda)))  # deterministic


@pytest.mark.parametrize('n,f', [(1, 1), (1000, 128), (100000, 256), (333, 1433)])
def test_colsum(cuda, n, f):
    x = torch.randn(n, f, generator=torch.Generator().manual_seed(n))
    assert rel_err(ops.colsum(x.to(cuda)), x.double().sum(0)) < FP32_TOL


def test_row_helpers(cuda):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(50, 20, generator=g)
    ids = torch.randperm(50, generator=g)[:13]
    assert torch.equal(ops.gather_rows(x.to(cuda), ids.to(cuda)).cpu(), x[ids])
    base = torch.randn(50, 20, generator=g)
    upd = torch.randn(13, 20, generator=g)
    got = ops.scatter_add_rows_(base.clone().to(cuda), ids.to(cuda), upd.to(cuda)).cpu()
    assert torch.equal(got, base.index_add(0, ids, upd))
    y = torch.randn(50, 20, generator=g)
    assert torch.equal(ops.relu_grad(x.to(cuda), y.to(cuda)).cpu(), x * (y > 0))


def test_tc_gemm_four_segments_sage_id_shape(cuda):
    """[x | mean] (W, W_id on centres): 4 K-segments into one TMEM tile, centres only in the first tile."""
    g = torch.Generator().manual_seed(5)
    n, k, f = 900, 20, 136
    x, m = torch.randn(n, k, generator=g), torch.randn(n, k, generator=g)
    w, wid = torch.randn(2 * k, f, generator=g), torch.randn(2 * k, f, generator=g)
    ids = torch.arange(0, 60, 3)
    cnt = ops.id_count(ids.to(cuda), n)
    d = lambda t: t.to(cuda)
    got = ops.id_gemm([(d(x), d(w[:k]), None), (d(m), d(w[k:]), None), (d(x), d(wid[:k]), cnt),
                       (d(m), d(wid[k:]), cnt)], n, f)
    z = torch.cat([x, m], 1).double()
    want = z @ w.double()
    want[ids] += z[ids] @ wid.double()
    assert rel_err(got, want) < FP32_TOL


@pytest.mark.parametrize('n,k,f,b_trans', [(148 * 128, 100, 128, False), (148 * 128 + 77, 128, 100, True),
                                           (60000, 36, 36, False), (33333, 4, 128, False), (40000, 64, 8, True),
                                           (148 * 128 * 3 + 1, 100, 128, False)])
def test_persistent_layer_gemm(cuda, gemm_mode, n, k, f, b_trans):
    """One unscaled segment, K <= 128, F <= 128, at least one row tile per SM: the persistent tcgen05 kernel (B image
    resident in shared memory, A slabs by TMA into a ring that never drains, two TMEM accumulators).  Strided A and
    out, bias + ReLU and the ReLU-mask epilogue; fp64 reference evaluated on the device."""
    g = torch.Generator(device=cuda).manual_seed(n + k + f)
    xa = torch.randn(n, k + 4, generator=g, device=cuda)
    x = xa[:, :k]                                              # row stride k + 4
    w = torch.randn((f, k) if b_trans else (k, f), generator=g, device=cuda)
    bias = torch.randn(f, generator=g, device=cuda)
    m = torch.randn(n, f, generator=g, device=cuda)
    wd = w.double().t() if b_trans else w.double()
    want = x.double() @ wd
    out = torch.full((n, f + 4), 7.0, device=cuda)
    got = ops.id_gemm([(x, w, None)], n, f, b_trans=b_trans, out=out[:, :f])
    assert rel_err(got, want) < FP32_TOL
    assert bool((out[:, f:] == 7.0).all())                      # nothing written outside the F columns
    got = ops.id_gemm([(x, w, None)], n, f, b_trans=b_trans, bias=bias, act=ops.ACT_RELU)
    assert rel_err(got, (want + bias.double()).relu()) < FP32_TOL
    got = ops.id_gemm([(x, w, None)], n, f, b_trans=b_trans, relu_mask=m)
    assert rel_err(got, want * (m > 0)) < FP32_TOL
    # row blocks are independent: the last rows equal a small-problem launch (per-tile kernel) of the same rows
    tail = ops.id_gemm([(x[-300:], w, None)], 300, f, b_trans=b_trans)
    full = ops.id_gemm([(x, w, None)], n, f, b_trans=b_trans)
    assert rel_err(full[-300:], tail) < 2e-6


@pytest.mark.parametrize('n,k,f', [(148 * 256, 100, 128), (148 * 256 + 333, 128, 100), (200000, 36, 36),
                                   (50000, 4, 128), (148 * 256 * 3 + 31, 64, 8)])
def test_persistent_weight_gradient(cuda, gemm_mode, n, k, f):
    """dW = X^T G for K, F <= 128 and at least one 256-row segment per SM: the persistent TN kernel (TMA slabs, A operand
    transposed into tensor memory, accumulator segments drained by TMA reduce-add into an L2-resident partial tile).
    Strided operands; fp64 reference evaluated on the device; two runs give the same bits."""
    g = torch.Generator(device=cuda).manual_seed(n + k + f)
    a = torch.randn(n, k + 4, generator=g, device=cuda)[:, :k]
    gr = torch.randn(n, f + 8, generator=g, device=cuda)[:, :f]
    got = ops.gemm_tn(a, gr)
    want = a.double().t() @ gr.double()
    assert rel_err(got, want) < FP32_TOL
    assert torch.equal(got, ops.gemm_tn(a, gr))
    # a column of ones in X turns its row of dW into the column sums of G
    ones = torch.ones(n, 4, device=cuda)
    cs = ops.gemm_tn(ones, gr)
    assert rel_err(cs[0], gr.double().sum(0)) < FP32_TOL


def test_persistent_gemms_are_deterministic_under_load(cuda, gemm_mode):
    """Regression test of the slot-release race (persistent kernels): at 600 K rows every CTA walks ~30 tiles
    with its rings full, which is where a refill landed under loads still in flight.  Repeated launches — bias + ReLU,
    ReLU mask, weight gradient — interleaved with a per-tile launch must reproduce the first results bit for bit, and the
    first results must be right."""
    if gemm_mode != 'tc':
        pytest.skip('tensor-core kernels only')
    n, k, f = 600_000, 100, 128
    gen = torch.Generator(device=cuda).manual_seed(11)
    x = torch.randn(n, k, generator=gen, device=cuda)
    g = torch.randn(n, f, generator=gen, device=cuda)
    w = torch.randn(k, f, generator=gen, device=cuda)
    b = torch.randn(f, generator=gen, device=cuda)
    small = torch.randn(3000, k, generator=gen, device=cuda)

    def launches():
        ops.id_gemm([(small, w, None)], 3000, f)
        return (ops.id_gemm([(x, w, None)], n, f, bias=b, act=ops.ACT_RELU),
                ops.id_gemm([(g, w, None)], n, k, b_trans=True, relu_mask=x),
                ops.gemm_tn(x, g))

    ref = [t.clone() for t in launches()]
    rows = slice(n - 70_000, n)                                   # the last tiles of every CTA
    assert rel_err(ref[0][rows], (x[rows].double() @ w.double() + b.double()).relu()) < FP32_TOL
    assert rel_err(ref[1][rows], (g[rows].double() @ w.double().t()) * (x[rows] > 0)) < FP32_TOL
    assert rel_err(ref[2], x.double().t() @ g.double()) < FP32_TOL
    for it in range(25):
        got = launches()
        for i, (a, r) in enumerate(zip(got, ref)):
            assert torch.equal(a, r), (it, i, int((a != r).sum()))
