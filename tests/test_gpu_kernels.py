"""GPU parity tests, kernel level (through the C ABI): dropout hash, GEMMs."""
import ctypes as C

import numpy as np
import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from music_generator_b200 import _lib
    return _lib.load()


def P(t):
    return C.c_void_p(t.data_ptr())


def test_dropout_masks_match_numpy_twin(lib):
    from music_generator_b200 import _lib
    for seed, site, rate, rows, F in ((7, 4, 0.5, 96, 64), (7, 9, 0.5, 50, 259), (11, 1, 0.2, 200, 3),
                                      (3, 5, 0.5, 48, 94), (5, 2, 0.3, 17, 16)):
        d = _lib.make_dropout(seed, site, rate)
        out = torch.empty(rows, F, device="cuda")
        _lib.check(lib.dj_dropout_mask_materialize(d, rows, F, P(out), None))
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), helpers.keep_mask(seed, site, rate, rows, F)), (seed, site)


@pytest.mark.parametrize("layout", ["nn", "nt", "tn", "tt"])
def test_gemm_simt_layouts(lib, layout):
    from music_generator_b200 import _lib
    g = torch.Generator().manual_seed(0)
    M, N, K = 150, 70, 333
    A = torch.randn(M, K, generator=g); B = torch.randn(K, N, generator=g); bias = torch.randn(N, generator=g)
    ref = (A.double() @ B.double() + bias.double()).numpy()
    Ad = (A if layout[0] == "n" else A.t().contiguous()).cuda()
    Bd = (B if layout[1] == "n" else B.t().contiguous()).cuda()
    a_sm, a_sk = (K, 1) if layout[0] == "n" else (1, M)
    b_sk, b_sn = (N, 1) if layout[1] == "n" else (1, K)
    Cd = torch.zeros(M, N, device="cuda")
    bd = bias.cuda()
    _lib.check(lib.dj_gemm_simt(P(Ad), 0, a_sm, a_sk, P(Bd), 0, b_sk, b_sn, P(Cd), N, P(bd), M, N, K, 0, 0, 0, None))
    torch.cuda.synchronize()
    assert helpers.rel_err(Cd.cpu().numpy(), ref) < 1e-5


def test_gemm_simt_splitk_shift_and_bf16(lib):
    from music_generator_b200 import _lib
    g = torch.Generator().manual_seed(1)
    # weight-gradient shape: C[64,96] += X[K,64]^T . dZ[K,96], K rows with a 1-step shift every 48 rows
    K, M, N = 48 * 200, 64, 96
    X = torch.randn(K, M, generator=g); dZ = torch.randn(K, N, generator=g)
    Xs = torch.zeros_like(X); Xs[1:] = X[:-1]; Xs[::48] = 0
    ref = (Xs.double().t() @ dZ.double()).numpy()
    Cd = torch.zeros(M, N, device="cuda")
    Xd, Zd = X.cuda(), dZ.cuda()     # keep device tensors alive: the launches are asynchronous
    _lib.check(lib.dj_gemm_simt(P(Xd), 0, 1, M, P(Zd), 0, N, 1, P(Cd), N, None, M, N, K, 1, 1, 48, None))
    torch.cuda.synchronize()
    assert helpers.rel_err(Cd.cpu().numpy(), ref) < 1e-4
    Xb, Zb = X.bfloat16(), dZ.bfloat16()
    ref = (Xb.double().t() @ Zb.double()).numpy()
    Cd.zero_()
    Xbd, Zbd = Xb.cuda(), Zb.cuda()
    _lib.check(lib.dj_gemm_simt(P(Xbd), 1, 1, M, P(Zbd), 1, N, 1, P(Cd), N, None, M, N, K, 1, 0, 0, None))
    torch.cuda.synchronize()
    assert helpers.rel_err(Cd.cpu().numpy(), ref) < 1e-4


@pytest.mark.parametrize("shape", [(256, 128, 64), (6144, 1024, 96), (6144, 512, 288), (1000, 94, 1024),
                                   (12288, 259, 512), (128 * 149, 1024, 256)])
def test_gate_gemm_tcgen05(lib, shape):
    """tcgen05/TMA GEMM against fp64 matmul of the same bf16-rounded operands."""
    from music_generator_b200 import _lib
    M, N, K = shape
    g = torch.Generator().manual_seed(2)
    lda = (K + 31) // 32 * 32
    A = torch.zeros(M, lda); A[:, :K] = torch.randn(M, K, generator=g)
    Bt = torch.zeros(N, lda); Bt[:, :K] = torch.randn(N, K, generator=g) * 0.1
    bias = torch.randn(N, generator=g)
    Ab, Bb = A.bfloat16(), Bt.bfloat16()
    ref = (Ab.double() @ Bb.double().t() + bias.double()).numpy()
    ldc = (N + 3) // 4 * 4
    Cd = torch.full((M, ldc), float("nan"), device="cuda")
    use_bias = N % 4 == 0
    Ad, Bd, bd = Ab.cuda(), Bb.cuda(), bias.cuda()   # keep alive across the asynchronous launch
    _lib.check(lib.dj_gate_gemm_bf16(P(Ad), lda, P(Bd), lda, P(Cd), ldc, P(bd) if use_bias else None, M, N, lda, None))
    torch.cuda.synchronize()
    got = Cd.cpu().numpy()[:, :N]
    if not use_bias:
        ref = ref - bias.double().numpy()
    assert np.isfinite(got).all()
    assert helpers.rel_err(got, ref) < 2e-5


@pytest.mark.parametrize("shape", [(6144, 128, 512), (98304, 256, 1024), (12288, 94, 1024), (6144 * 3, 259, 512)])
def test_wgrad_gemm_tcgen05(lib, shape):
    """C[Ka,Nb] += A[M,Ka]^T.B[M,Nb] on tcgen05 (MN-major operands, split over M)."""
    from music_generator_b200 import _lib
    M, Ka, Nb = shape
    g = torch.Generator().manual_seed(3)
    lda = (Ka + 31) // 32 * 32
    A = torch.zeros(M, lda); A[:, :Ka] = torch.randn(M, Ka, generator=g)
    B = torch.randn(M, Nb, generator=g) * 0.1
    Ab, Bb = A.bfloat16(), B.bfloat16()
    C0 = torch.randn(Ka, Nb, generator=g)
    ref = (C0.double() + Ab[:, :Ka].double().t() @ Bb.double()).numpy()
    Ad, Bd, Cd = Ab.cuda(), Bb.cuda(), C0.cuda()
    _lib.check(lib.dj_wgrad_gemm_bf16(P(Ad), lda, P(Bd), Nb, P(Cd), Nb, Ka, Nb, M, None))
    torch.cuda.synchronize()
    assert helpers.rel_err(Cd.cpu().numpy(), ref) < 5e-5
