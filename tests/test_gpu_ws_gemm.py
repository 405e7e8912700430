"""GPU tests of the warp-specialised TMA/DMMA GEMM (csrc/dgemm_ws.cuh) through qcp2_dgemm_ex: every CTA tile x every
operand layout x edge shapes, forced split-K (in-kernel ordered reduction), strided batches and the transposed
problem, against numpy / torch fp64 matmul.  FP64: tolerance 1e-12 * sqrt(K) relative to unit-variance inputs."""
import numpy as np
import pytest

import qcp2_b200

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

TILES = {0: (80, 256), 1: (48, 256), 2: (16, 256), 3: (64, 64), 4: (112, 128), 5: (16, 192), 6: (32, 256), 7: (96, 128)}


@pytest.fixture(scope="module")
def dctx():
    c = qcp2_b200.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    c.set_stream(stream.cuda_stream)
    c.set_pointer_mode(qcp2_b200.DEVICE)
    yield c
    c.close()


def _operands(M, N, K, tA, tB, batch, seed, pad=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    f = dict(dtype=torch.float64, device="cuda")
    sa = (K, M) if tA else (M, K)
    sb = (N, K) if tB else (K, N)
    # even leading dimensions (16-byte rows) are what the TMA path needs; `pad` adds unused columns
    lda = sa[1] + (sa[1] & 1) + pad
    ldb = sb[1] + (sb[1] & 1) + pad
    ldc = N + (N & 1) + pad
    A = torch.randn((batch, sa[0], lda), generator=g, **f)
    B = torch.randn((batch, sb[0], ldb), generator=g, **f)
    Cm = torch.randn((batch, M, ldc), generator=g, **f)
    return A, B, Cm, lda, ldb, ldc, sa, sb


def _check(dctx, M, N, K, tA, tB, tile, split=0, batch=1, alpha=0.75, beta=-0.5, seed=1, pad=0):
    A, B, Cm, lda, ldb, ldc, sa, sb = _operands(M, N, K, tA, tB, batch, seed, pad)
    C0 = Cm.clone()
    ws0, lg0 = dctx.gemm_counts()
    dctx.dgemm_ex(tA, tB, M, N, K, alpha, A, lda, B, ldb, beta, Cm, ldc, batch=batch, strideA=A.stride(0), strideB=B.stride(0),
                  strideC=Cm.stride(0), tile=tile, k_splits=split)
    ws1, lg1 = dctx.gemm_counts()
    assert ws1 == ws0 + 1 and lg1 == lg0, "the product did not run on the warp-specialised kernel"
    a = A[:, :, :sa[1]].transpose(1, 2) if tA else A[:, :, :sa[1]]
    b = B[:, :, :sb[1]].transpose(1, 2) if tB else B[:, :, :sb[1]]
    want = alpha * torch.matmul(a, b) + beta * C0[:, :, :N]
    err = float((Cm[:, :, :N] - want).abs().max())
    assert err < 1e-12 * max(1.0, K ** 0.5) * 4, (M, N, K, tA, tB, tile, split, err)
    # padding columns of C are never written
    assert torch.equal(Cm[:, :, N:], C0[:, :, N:])


@pytest.mark.parametrize("tile", sorted(TILES))
@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_every_tile_and_layout_with_edges(dctx, tile, tA, tB):
    bm, bn = TILES[tile]
    m_cap = bm if tile in (1, 2, 5, 6) else 3 * bm  # the skinny tiles are chosen for M <= 48 / 16 only
    for (M, N, K) in [(min(m_cap, bm + 3), bn + 5, 37), (min(m_cap, 2 * bm), 2 * bn, 64), (min(m_cap, 7), 9, 5), (min(m_cap, bm - 1), 3 * bn - 2, 16 * 9 + 1)]:
        _check(dctx, M, N, K, tA, tB, tile, seed=M + N + K)


@pytest.mark.parametrize("tile", sorted(TILES))
def test_transposed_problem(dctx, tile):
    bm, bn = TILES[tile]
    for tA, tB in [(0, 0), (1, 1), (0, 1)]:
        _check(dctx, 2 * bn + 3, min(bm, 11), 50, tA, tB, tile + 16, seed=tile)


@pytest.mark.parametrize("tile,split", [(0, 3), (1, 8), (2, 5), (3, 37), (4, 11), (3, 293), (5, 148)])
@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0)])
def test_in_kernel_split_k_is_exact_and_repeatable(dctx, tile, split, tA, tB):
    bm, bn = TILES[tile]
    M, N, K = min(bm, 45), bn + 9, 16 * split * 5 + 7
    _check(dctx, M, N, K, tA, tB, tile, split=split, seed=split)
    # same launch twice: bit-identical (the partials are summed in split order whichever CTA arrives last), and the
    # tile counters went back to zero
    A, B, Cm, lda, ldb, ldc, _, _ = _operands(M, N, K, tA, tB, 1, 5)
    C1, C2 = torch.zeros_like(Cm), torch.zeros_like(Cm)
    for out in (C1, C2):
        dctx.dgemm_ex(tA, tB, M, N, K, 1.0, A, lda, B, ldb, 0.0, out, ldc, tile=tile, k_splits=split)
    assert torch.equal(C1, C2)


@pytest.mark.parametrize("tile", [0, 2, 3])
def test_strided_batch_and_shared_operand(dctx, tile):
    _check(dctx, 33, 300, 70, 0, 0, tile, batch=5, seed=3)
    _check(dctx, 16, 130, 45, 1, 1, tile, batch=3, split=2, seed=4)
    # B shared by every batch member (stride 0), as in cc.cpp:237 (W_abef[a] slabs)
    A, B, Cm, lda, ldb, ldc, _, _ = _operands(10, 200, 30, 0, 0, 4, 9)
    dctx.dgemm_ex(0, 0, 10, 200, 30, 1.0, A, lda, B[0], ldb, 0.0, Cm, ldc, batch=4, strideA=A.stride(0), strideB=0,
                  strideC=Cm.stride(0), tile=tile)
    want = torch.matmul(A[:, :, :30], B[0][:, :200])
    assert float((Cm[:, :, :200] - want).abs().max()) < 1e-11


def test_automatic_choice_on_ccsd_shapes(dctx):
    """Shapes of one CCSD iteration at o=10, v=162 (scaled down in the long dimension): whatever tile / split the cost
    model picks must reproduce torch.matmul."""
    for (M, N, K, tA, tB) in [(46, 1304, 1304, 0, 1), (100, 1620, 2624, 0, 1), (324, 324, 1620, 0, 0), (10, 26244, 162, 0, 1),
                              (26244, 10, 162, 0, 0), (100, 100, 26244, 0, 1), (162, 162, 16200, 1, 0), (10, 162, 262440 // 4, 0, 1)]:
        _check(dctx, M, N, K, tA, tB, -1, seed=K)
