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


def _gates16(G):
    """fp32 activated gates -> the IEEE-half copy the tensor-core scans exchange: round to nearest, but a hard-sigmoid
    value strictly inside (0, 1) stays strictly inside (columns 4u+{0,1,3}; column 4u+2 is the tanh gate)."""
    H = G.half()
    sig = torch.ones(G.shape[-1], dtype=torch.bool, device=G.device)
    sig[2::4] = False
    one, tiny = torch.tensor(0x3BFF, dtype=torch.int16, device=G.device).view(torch.float16), 5.9604645e-08
    H = torch.where(sig & (G < 1) & (H == 1), one, H)
    H = torch.where(sig & (G > 0) & (H == 0), torch.full_like(H, tiny), H)
    return H.contiguous()


def _split16(x, dt):
    hi = x.to(dt)
    return hi, (x - hi.float()).to(dt)


@pytest.mark.parametrize("shape", [(6144, 1024, 96), (12288, 512, 288), (128 * 149, 1024, 256), (1000, 512, 128)])
def test_gate_gemm_split_is_fp32_grade(lib, shape):
    """dj_gate_gemm_16 with residual operands: A.B + A_lo.B + A.B_lo in three tcgen05 passes on one TMEM accumulator
    must reproduce the fp64 product of the UNROUNDED fp32 operands to ~2^-16 (single bf16 operands: ~2^-9)."""
    from music_generator_b200 import _lib
    M, N, K = shape
    g = torch.Generator().manual_seed(12)
    A = torch.randn(M, K, generator=g)
    Bt = torch.randn(N, K, generator=g) * 0.1
    bias = torch.randn(N, generator=g)
    ref = (A.double() @ Bt.double().t() + bias.double()).numpy()
    Ah, Al = _split16(A, torch.bfloat16)
    Bh, Bl = _split16(Bt, torch.bfloat16)
    Cd = torch.full((M, N), float("nan"), device="cuda")
    dev = [t.cuda() for t in (Ah, Al, Bh, Bl, bias)]
    _lib.check(lib.dj_gate_gemm_16(P(dev[0]), P(dev[1]), 1, K, P(dev[2]), P(dev[3]), 1, K, P(Cd), N, P(dev[4]), M, N, K, None))
    torch.cuda.synchronize()
    got = Cd.cpu().numpy()
    assert np.isfinite(got).all()
    err_split = helpers.rel_err(got, ref)
    _lib.check(lib.dj_gate_gemm_16(P(dev[0]), None, 1, K, P(dev[2]), None, 1, K, P(Cd), N, P(dev[4]), M, N, K, None))
    torch.cuda.synchronize()
    err_single = helpers.rel_err(Cd.cpu().numpy(), ref)
    assert err_split < 2e-5 and err_single > 20 * err_split, (err_split, err_single)


def test_gemm_rejects_mixed_operand_formats(lib):
    """tcgen05 kind::f16 with one half and one bf16 operand raises an illegal-instruction fault on B200 (measured in
    round 2), so the ABI refuses the combination instead of launching it."""
    a = torch.zeros(128, 64, device="cuda", dtype=torch.float16)
    c = torch.zeros(128, 128, device="cuda")
    assert lib.dj_gate_gemm_16(P(a), None, 2, 64, P(a), None, 1, 64, P(c), 128, None, 128, 128, 64, None) == -1
    assert b"cannot mix" in lib.dj_last_error()
    assert lib.dj_wgrad_gemm_16(P(a), 1, 64, P(a), 2, 64, P(c), 128, 64, 64, 128, None) == -1


def test_half_to_bf16_inplace(lib):
    from music_generator_b200 import _lib
    x = (torch.randn(4096 * 8, generator=torch.Generator().manual_seed(3)) * 0.7).half()
    buf = x.cuda()
    _lib.check(lib.dj_half_to_bf16_inplace(P(buf), buf.numel(), None))
    torch.cuda.synchronize()
    assert torch.equal(buf.view(torch.bfloat16).float().cpu(), x.float().bfloat16().float())


@pytest.mark.parametrize("fmts", [(2, 2), (1, 1)])
def test_gemm_operand_formats_half(lib, fmts):
    """kind::f16 with both operands IEEE half (the forward recurrence's format) or both bf16."""
    from music_generator_b200 import _lib
    dts = {1: torch.bfloat16, 2: torch.float16}
    g = torch.Generator().manual_seed(13)
    M, N, K = 6144, 512, 256
    A = torch.randn(M, K, generator=g).to(dts[fmts[0]])
    Bt = (torch.randn(N, K, generator=g) * 0.1).to(dts[fmts[1]])
    ref = (A.double() @ Bt.double().t()).numpy()
    Ad, Bd = A.cuda(), Bt.cuda()
    Cd = torch.full((M, N), float("nan"), device="cuda")
    _lib.check(lib.dj_gate_gemm_16(P(Ad), None, fmts[0], K, P(Bd), None, fmts[1], K, P(Cd), N, None, M, N, K, None))
    torch.cuda.synchronize()
    assert helpers.rel_err(Cd.cpu().numpy(), ref) < 2e-5
    # weight-gradient kernel (MN-major operands): C[K,N] += A[M,K]^T.B[M,N]
    Bm = (torch.randn(M, N, generator=g) * 0.1).to(dts[fmts[1]])
    ref = (A.double().t() @ Bm.double()).numpy()
    Bmd = Bm.cuda()
    Cw = torch.zeros(K, N, device="cuda")
    _lib.check(lib.dj_wgrad_gemm_16(P(Ad), fmts[0], K, P(Bmd), fmts[1], N, P(Cw), N, K, N, M, None))
    torch.cuda.synchronize()
    assert helpers.rel_err(Cw.cpu().numpy(), ref) < 5e-5


@pytest.mark.parametrize("shape", [(6144, 1024, 96), (6144, 1024, 256), (1536 * 128, 1024, 256)])
def test_gate_gemm_half_split_scaled_is_fp32_grade(lib, shape):
    """The generation path's projection: IEEE half hi+lo operands, weights pre-scaled by 2^10 (undone by out_scale):
    fp32-grade (the operands carry ~22 bits; what is left is the tensor core's truncating fp32 accumulation)."""
    from music_generator_b200 import _lib
    M, N, K = shape
    g = torch.Generator().manual_seed(14)
    A = torch.randn(M, K, generator=g) * torch.rand(M, 1, generator=g)          # rows of very different sizes
    Bt = torch.randn(N, K, generator=g) * 0.07
    bias = torch.randn(N, generator=g)
    ref = (A.double() @ Bt.double().t() + bias.double()).numpy()
    Ah, Al = _split16(A, torch.float16)
    Bh, Bl = _split16(Bt * 1024.0, torch.float16)
    Cd = torch.full((M, N), float("nan"), device="cuda")
    dev = [t.cuda() for t in (Ah, Al, Bh, Bl, bias)]
    _lib.check(lib.dj_gate_gemm_16s(P(dev[0]), P(dev[1]), 2, K, P(dev[2]), P(dev[3]), 2, K, P(Cd), N, P(dev[4]),
                                    1.0 / 1024.0, M, N, K, None))
    torch.cuda.synchronize()
    got = Cd.cpu().numpy()
    err = float(np.abs(got - ref).max())
    ref32 = float(np.abs((A @ Bt.t() + bias).numpy() - ref).max())            # torch's fp32 CPU matmul, for scale
    print(f"half-split GEMM {shape}: max abs err {err:.2e} (fp32 CPU matmul: {ref32:.2e})")
    # not 2^-23: tcgen05 accumulates the fp32 partial sums with truncation, a bias of up to an ulp per MMA (48 - 144
    # MMAs deep here), i.e. a few 1e-6 relative -- measured 3-4x the error of an fp32 FMA-chain matmul
    assert np.isfinite(got).all() and err < 4e-5 and err < 8 * ref32


@pytest.mark.parametrize("B,T", [(1, 128), (3, 16), (32, 128), (40, 8)])
def test_lstm_scan_tcgen05_inference_is_fp32_grade(lib, B, T):
    """dj_lstm_scan_tc_infer (generation window: h and U as half hi+lo, three MMA passes) against the fp32 CUDA-core
    recurrence: the two must agree like two fp32 implementations do (1e-6), and Z must be left untouched."""
    from music_generator_b200 import _lib
    g = torch.Generator().manual_seed(15)
    U, M = 256, B * T * 48
    Z0 = torch.randn(M, 4 * U, generator=g)
    Uw = torch.randn(U, 4 * U, generator=g) * 0.06
    S, steps, m = B * 48, T, (48, T * 48, 1, 48)
    Zr, Zt, Ud = Z0.cuda(), Z0.cuda(), Uw.cuda()
    Ut, Ut_lo = _split16(Uw.t().contiguous() * 1024.0, torch.float16)
    Ut, Ut_lo = Ut.cuda(), Ut_lo.cuda()
    hr, cr = torch.empty(M, U, device="cuda"), torch.empty(M, U, device="cuda")
    ht = torch.zeros(M, U, device="cuda")
    hhi = torch.zeros(M, U, device="cuda", dtype=torch.float16)
    hlo = torch.zeros(M, U, device="cuda", dtype=torch.float16)
    _lib.check(lib.dj_lstm_scan_fwd(P(Zr), P(hr), P(cr), None, P(Ud), S, steps, U, *m, 1, None))
    _lib.check(lib.dj_lstm_scan_tc_infer(P(Zt), P(ht), P(hhi), P(hlo), P(Ut), P(Ut_lo), 1.0 / 1024.0, S, steps, U, *m, 1, None))
    torch.cuda.synchronize()
    dh = float((ht - hr).abs().max())
    print(f"scan_tc_infer[B={B},T={T}]: max|dh| {dh:.2e}")
    assert torch.isfinite(ht).all() and dh < 2e-6
    assert torch.equal(Zt.cpu(), Z0)


@pytest.mark.parametrize("axis,B,T", [("time", 1, 12), ("time", 5, 7), ("time", 8, 6), ("note", 1, 5), ("note", 13, 40)])
def test_lstm_scan_fp32_matches_recurrence(lib, axis, B, T):
    """fp32 CUDA-core forward recurrence (both tilings: 16-sequence tiles / one unit per thread for few
    sequences, 32-sequence tiles otherwise; ragged last tiles) against a float64 restatement of the
    Keras LSTM step (gate order i,f,c,o, hard_sigmoid, zero initial state)."""
    from music_generator_b200 import _lib
    g = torch.Generator().manual_seed(11)
    U = 256 if axis == "time" else 128
    M = B * T * 48
    Z0 = torch.randn(M, 4 * U, generator=g)            # gate-interleaved columns: col = 4*unit + gate
    Uw = torch.randn(U, 4 * U, generator=g) * 0.06
    if axis == "time":
        S, steps, m = B * 48, T, (48, T * 48, 1, 48)
        rows = torch.arange(M).view(B, T, 48).permute(0, 2, 1).reshape(S, steps)      # [seq, step] -> row
    else:
        S, steps, m = B * T, 48, (1, 48, 0, 1)
        rows = torch.arange(M).view(S, steps)
    Zd, Ud = Z0.cuda(), Uw.cuda()
    h, c = torch.empty(M, U, device="cuda"), torch.empty(M, U, device="cuda")
    _lib.check(lib.dj_lstm_scan_fwd(P(Zd), P(h), P(c), None, P(Ud), S, steps, U, *m, 1, None))
    torch.cuda.synchronize()
    Z64, U64 = Z0.double(), Uw.double()
    hs, cs = torch.zeros(S, U, dtype=torch.float64), torch.zeros(S, U, dtype=torch.float64)
    href, cref = torch.zeros(M, U, dtype=torch.float64), torch.zeros(M, U, dtype=torch.float64)
    for t in range(steps):
        z = (Z64[rows[:, t]] + hs @ U64).view(S, U, 4)
        i, f, o = [(0.2 * z[..., k] + 0.5).clamp(0, 1) for k in (0, 1, 3)]
        cs = f * cs + i * torch.tanh(z[..., 2])
        hs = o * torch.tanh(cs)
        href[rows[:, t]], cref[rows[:, t]] = hs, cs
    assert float((h.cpu().double() - href).abs().max()) < 2e-5
    assert float((c.cpu().double() - cref).abs().max()) < 2e-5


@pytest.mark.parametrize("mode", ["bf16", "half_split"])
@pytest.mark.parametrize("axis,B,T", [("time", 3, 8), ("note", 2, 32), ("time", 20, 4), ("note", 40, 128),
                                      ("note256", 2, 32), ("time", 40, 24)])
def test_lstm_scan_tcgen05_matches_fp32_scan(lib, axis, B, T, mode):
    """Tensor-core recurrence against the fp32 CUDA-core recurrence (itself checked against a float64 restatement of
    the Keras step above) on the same pre-activations.  `bf16`: h and U one bf16 each (weights chosen
    bf16-representable, so the difference is the rounding of h).  `half_split` (the training default): h in IEEE half,
    U as half hi + lo in two MMA passes, ARBITRARY fp32 weights -- must track the fp32 recurrence 10x closer."""
    from music_generator_b200 import _lib
    g = torch.Generator().manual_seed(5)
    U = 256 if axis in ("time", "note256") else 128
    M = B * T * 48
    Z0 = torch.randn(M, 4 * U, generator=g)
    Uw = torch.randn(U, 4 * U, generator=g) * 0.06
    if mode == "bf16":
        Uw = Uw.bfloat16().float()                                          # bf16-representable weights
    if axis == "time":
        S, steps, m = B * 48, T, (48, T * 48, 1, 48)
    else:
        S, steps, m = B * T, 48, (1, 48, 0, 1)
    Zr, Zt, Ud = Z0.cuda(), Z0.cuda(), Uw.cuda()
    hdt = torch.bfloat16 if mode == "bf16" else torch.float16
    Ut, Ut_lo = _split16(Uw.t().contiguous(), hdt)
    Ut, Ut_lo = Ut.cuda(), Ut_lo.cuda()
    hr, cr = torch.empty(M, U, device="cuda"), torch.empty(M, U, device="cuda")
    ht, ct = torch.zeros(M, U, device="cuda"), torch.zeros(M, U, device="cuda")
    hp = torch.full((M, U), 7.0, device="cuda").to(hdt)
    _lib.check(lib.dj_lstm_scan_fwd(P(Zr), P(hr), P(cr), None, P(Ud), S, steps, U, *m, 1, None))
    G16 = torch.zeros(M, 4 * U, device="cuda", dtype=torch.float16)
    _lib.check(lib.dj_lstm_scan_tc_fwd(P(Zt), P(G16), P(ht), P(ct), P(hp), P(Ut), P(Ut_lo) if mode != "bf16" else None,
                                       1 if mode == "bf16" else 2, S, steps, U, *m, 1, None))
    torch.cuda.synchronize()
    dh = (ht - hr).abs()
    assert torch.isfinite(ht).all()
    tol = dict(bf16=(0.03, 2e-3, 0.06), half_split=(3e-3, 1e-4, 6e-3))[mode]
    print(f"scan_tc_fwd[{axis},{B},{T},{mode}]: max|dh| {float(dh.max()):.2e} mean {float(dh.mean()):.2e} "
          f"max|dc| {float((ct - cr).abs().max()):.2e}")
    assert float(dh.max()) < tol[0] and float(dh.mean()) < tol[1], (float(dh.max()), float(dh.mean()))
    assert float((ct - cr).abs().max()) < tol[2]
    assert torch.equal(Zt.cpu(), Z0)                       # the pre-activations are read only
    # activated gates saved as IEEE half for the reverse scan: the fp32 scan's gates (saved in place), to half precision
    assert float((G16.float() - Zr).abs().mean()) < tol[1] + 1.5e-4
    if mode == "half_split":
        # the hard-sigmoid indicator 0 < a < 1 survives the rounding to half: gates the fp32 scan has within 2e-4 of
        # an end (closer than half's spacing there, but further than the two scans differ) must not sit ON the end
        g = G16.float()
        sig = torch.ones(4 * U, dtype=torch.bool, device="cuda")
        sig[2::4] = False
        near1 = sig & (Zr < 1 - 2e-5) & (Zr > 1 - 2e-4)
        near0 = sig & (Zr > 2e-5) & (Zr < 2e-4)
        assert bool((g[near1] < 1).all()) and bool((g[near0] > 0).all())
    # hprev = bf16(h) shifted by one step, zero at step 0
    h4, p4 = ht.view(B, T, 48, U), hp.float().view(B, T, 48, U)
    want = torch.zeros_like(h4)
    if axis == "time":
        want[:, 1:] = h4[:, :-1]
    else:
        want[:, :, 1:] = h4[:, :, :-1]
    assert torch.equal(p4, want.to(hdt).float())


def _keras_lstm_f64(Z0, Uw, rows, S, steps, U):
    """float64 restatement of the Keras LSTM step (gate-interleaved columns i,f,c,o, hard_sigmoid, zero
    initial state) over rows[seq, step]."""
    Z64, U64 = Z0.double(), Uw.double()
    hs, cs = torch.zeros(S, U, dtype=torch.float64), torch.zeros(S, U, dtype=torch.float64)
    href = torch.zeros(Z0.shape[0], U, dtype=torch.float64)
    for t in range(steps):
        z = (Z64[rows[:, t]] + hs @ U64).view(S, U, 4)
        i, f, o = [(0.2 * z[..., k] + 0.5).clamp(0, 1) for k in (0, 1, 3)]
        cs = f * cs + i * torch.tanh(z[..., 2])
        hs = o * torch.tanh(cs)
        href[rows[:, t]] = hs
    return href


@pytest.mark.parametrize("axis,B,T", [("time", 2, 24), ("note", 2, 32), ("time", 34, 6)])
def test_lstm_scan_tcgen05_matches_float64_recurrence(lib, axis, B, T):
    """The training recurrence on the tensor cores (h in IEEE half, U as half hi + lo) DIRECTLY against the float64
    restatement of the Keras step, not through the fp32 kernel: the error is the half rounding of h, 1e-4 class."""
    from music_generator_b200 import _lib
    g = torch.Generator().manual_seed(21)
    U = 256 if axis == "time" else 128
    M = B * T * 48
    Z0 = torch.randn(M, 4 * U, generator=g)
    Uw = torch.randn(U, 4 * U, generator=g) * 0.06
    if axis == "time":
        S, steps, m = B * 48, T, (48, T * 48, 1, 48)
        rows = torch.arange(M).view(B, T, 48).permute(0, 2, 1).reshape(S, steps)
    else:
        S, steps, m = B * T, 48, (1, 48, 0, 1)
        rows = torch.arange(M).view(S, steps)
    Ut, Ut_lo = _split16(Uw.t().contiguous(), torch.float16)
    Zt, Ut, Ut_lo = Z0.cuda(), Ut.cuda(), Ut_lo.cuda()
    ht, ct = torch.zeros(M, U, device="cuda"), torch.zeros(M, U, device="cuda")
    hp = torch.zeros(M, U, device="cuda", dtype=torch.float16)
    G16 = torch.zeros(M, 4 * U, device="cuda", dtype=torch.float16)
    _lib.check(lib.dj_lstm_scan_tc_fwd(P(Zt), P(G16), P(ht), P(ct), P(hp), P(Ut), P(Ut_lo), 2, S, steps, U, *m, 1, None))
    torch.cuda.synchronize()
    href = _keras_lstm_f64(Z0, Uw, rows, S, steps, U)
    err = (ht.cpu().double() - href).abs()
    print(f"scan_tc_fwd vs float64 [{axis},{B},{T}]: max {float(err.max()):.2e} mean {float(err.mean()):.2e}")
    assert float(err.max()) < 1e-3 and float(err.mean()) < 3e-5


def test_lstm_scan_tcgen05_inference_matches_float64_recurrence(lib):
    """The generation window's recurrence (h and U as half hi + lo, three MMA passes) directly against float64."""
    from music_generator_b200 import _lib
    g = torch.Generator().manual_seed(22)
    B, T, U = 1, 128, 256
    M = B * T * 48
    Z0 = torch.randn(M, 4 * U, generator=g)
    Uw = torch.randn(U, 4 * U, generator=g) * 0.06
    S, steps, m = B * 48, T, (48, T * 48, 1, 48)
    rows = torch.arange(M).view(B, T, 48).permute(0, 2, 1).reshape(S, steps)
    scale = 1024.0
    Ut, Ut_lo = _split16((Uw.t().contiguous() * scale), torch.float16)
    Zt, Ut, Ut_lo = Z0.cuda(), Ut.cuda(), Ut_lo.cuda()
    ht = torch.zeros(M, U, device="cuda")
    hhi = torch.zeros(M + 48, U, device="cuda", dtype=torch.float16)
    hlo = torch.zeros(M + 48, U, device="cuda", dtype=torch.float16)
    _lib.check(lib.dj_lstm_scan_tc_infer(P(Zt), P(ht), P(hhi), P(hlo), P(Ut), P(Ut_lo), 1.0 / scale, S, steps, U, *m, 1, None))
    torch.cuda.synchronize()
    href = _keras_lstm_f64(Z0, Uw, rows, S, steps, U)
    err = (ht.cpu().double() - href).abs()
    print(f"scan_tc_infer vs float64: max {float(err.max()):.2e} mean {float(err.mean()):.2e}")
    assert float(err.max()) < 5e-6


@pytest.mark.parametrize("axis,B,T", [("time", 3, 8), ("note", 2, 32), ("note", 40, 128), ("note256", 2, 32),
                                      ("time", 40, 6), ("time", 64, 5),   # > 33 tiles: 96-sequence tiles, shared staging
                                      ("time", 35, 4),                     # odd batch: two waves of 48-sequence tiles
                                      ("time", 2, 8), ("time", 68, 3)])    # one tile pair; more clusters than are resident
def test_lstm_scan_bwd_tcgen05_matches_fp32_scan(lib, axis, B, T):
    """Reverse scan on tcgen05 (bf16 dz.U^T) against the fp32 CUDA-core reverse scan."""
    from music_generator_b200 import _lib
    g = torch.Generator().manual_seed(6)
    U = 256 if axis in ("time", "note256") else 128
    M = B * T * 48
    Z0 = torch.randn(M, 4 * U, generator=g)
    Uw = (torch.randn(U, 4 * U, generator=g) * 0.06).bfloat16().float()
    if axis == "time":
        S, steps, m = B * 48, T, (48, T * 48, 1, 48)
    else:
        S, steps, m = B * T, 48, (1, 48, 0, 1)
    Z, Ud = Z0.cuda(), Uw.cuda()
    h, c = torch.empty(M, U, device="cuda"), torch.empty(M, U, device="cuda")
    _lib.check(lib.dj_lstm_scan_fwd(P(Z), P(h), P(c), None, P(Ud), S, steps, U, *m, 1, None))
    ld = U + 32
    dY = (torch.randn(M, ld, generator=g) * 0.1).cuda()
    d = _lib.make_dropout(11, 6, 0.5)
    dZr = torch.zeros(M, 4 * U, device="cuda")
    dbr = torch.zeros(4 * U, device="cuda")
    _lib.check(lib.dj_lstm_scan_bwd(P(Z), P(c), P(dY), ld, d, P(Ud), P(dZr), 0, P(dbr), S, steps, U, *m, 1, None))
    Un = Uw.bfloat16().cuda()
    dZt = torch.zeros(M, 4 * U, device="cuda").bfloat16()
    dbt = torch.zeros(4 * U, device="cuda")
    G16 = _gates16(Z)
    _lib.check(lib.dj_lstm_scan_tc_bwd(P(G16), P(c), P(dY), ld, d, P(Un), P(dZt), P(dbt), S, steps, U, *m, 1, None))
    torch.cuda.synchronize()
    a, b = dZt.float().cpu().numpy(), dZr.cpu().numpy()
    assert np.isfinite(a).all()
    scale = np.abs(b).max()
    assert np.abs(a - b).max() / scale < 0.03 and np.abs(a - b).mean() / np.abs(b).mean() < 0.02
    assert helpers.rel_err(dbt.cpu().numpy(), dbr.cpu().numpy()) < 0.02


@pytest.mark.parametrize("G", [1, 2])
def test_generation_two_layer_scan_matches_float64_and_stays_in_bounds(lib, G):
    """dj_lstm_scan_tc_gen2 (both time-axis layers of a generation window in one launch; G = 1: three roles, G = 2:
    two) against the float64 two-layer recurrence  z1_t = h0_t.W1 + c1 + h1_{t-1}.U1  on random weights, and a bounds
    check: every buffer carries a guard region of sentinels behind its documented size that must come back intact."""
    from music_generator_b200 import _lib
    g = torch.Generator().manual_seed(40 + G)
    U, T = 256, 24
    S, M = G * 48, G * T * 48
    Z0 = torch.randn(M, 4 * U, generator=g)
    U0 = torch.randn(U, 4 * U, generator=g) * 0.06
    U1 = torch.randn(U, 4 * U, generator=g) * 0.06
    W1 = torch.randn(U, 4 * U, generator=g) * 0.06
    c1 = torch.randn(G, 4 * U, generator=g) * 0.5
    rows = torch.arange(M).view(G, T, 48).permute(0, 2, 1).reshape(S, T)
    # float64 reference
    h0 = _keras_lstm_f64(Z0, U0, rows, S, T, U)
    Z1 = h0 @ W1.double() + c1.double().repeat_interleave(T * 48, dim=0)
    h1 = _keras_lstm_f64(Z1, U1, rows, S, T, U)
    # device
    scale, GUARD, SENT = 1024.0, 64, 12345.0
    def split_t(Wm):
        hi, lo = _split16(Wm.t().contiguous() * scale, torch.float16)
        return hi.cuda(), lo.cuda()
    Ut0, Ut0lo = split_t(U0); Ut1, Ut1lo = split_t(U1); Wt1, Wt1lo = split_t(W1)
    def guarded(n_rows, cols, dtype):
        t = torch.full((n_rows + GUARD, cols), SENT, dtype=dtype, device="cuda")
        return t
    hbuf = [guarded(M + 48, U, torch.float16) for _ in range(4)]          # h0_hi, h0_lo, h1_hi, h1_lo
    h1out = guarded(M, U, torch.float32)
    Z1d = guarded(M, 4 * U, torch.float32)
    flags = torch.full((2 * (S // 16) + 16,), 77, dtype=torch.int32, device="cuda")
    Zd, c1d = Z0.cuda(), c1.cuda()
    _lib.check(lib.dj_lstm_scan_tc_gen2(P(Zd), P(Z1d), P(c1d), P(h1out), P(hbuf[0]), P(hbuf[1]), P(hbuf[2]), P(hbuf[3]),
                                        P(Ut0), P(Ut0lo), P(Ut1), P(Ut1lo), P(Wt1), P(Wt1lo), 1.0 / scale, P(flags),
                                        S, T, 1, None))
    torch.cuda.synchronize()
    last = rows[:, T - 1]
    err = (h1out[:M].cpu().double()[last] - h1[last]).abs()
    print(f"scan_tc_gen2 G={G}: max |h1 - float64| at the last step {float(err.max()):.2e}")
    assert float(err.max()) < 1e-5
    for t in hbuf:
        assert bool((t[M + 48:] == SENT).all())                            # nothing behind the spare timestep
    assert bool((hbuf[2][M:] == SENT).all()) and bool((hbuf[3][M:] == SENT).all())   # layer 1 never uses its spare rows
    assert bool((h1out[M:] == SENT).all()) and bool((Z1d[M:] == SENT).all())
    assert bool((flags[2 * (S // 16):] == 77).all())
    assert torch.equal(Zd.cpu(), Z0)                                       # the pre-activations are read only


@pytest.mark.parametrize("axis,B,T", [("time", 2, 12), ("note", 2, 32)])
def test_lstm_scan_bwd_tcgen05_matches_float64_autograd(lib, axis, B, T):
    """The reverse scan on the tensor cores against AUTOGRAD through the float64 restatement of the Keras step (not
    through the repo's fp32 reverse scan): dZ = d(sum dY*mask*h)/dZ and its column sum db, with the dropout mask of
    the layer output replayed by the NumPy twin.  Operands of the recurrent product are bf16: 1 % class."""
    from music_generator_b200 import _lib
    g = torch.Generator().manual_seed(31)
    U = 256 if axis == "time" else 128
    M = B * T * 48
    Z0 = torch.randn(M, 4 * U, generator=g)
    Uw = (torch.randn(U, 4 * U, generator=g) * 0.06).bfloat16().float()       # bf16-representable recurrent weights
    if axis == "time":
        S, steps, m = B * 48, T, (48, T * 48, 1, 48)
        rows = torch.arange(M).view(B, T, 48).permute(0, 2, 1).reshape(S, steps)
    else:
        S, steps, m = B * T, 48, (1, 48, 0, 1)
        rows = torch.arange(M).view(S, steps)
    ld = U + 32
    dY = torch.randn(M, ld, generator=g) * 0.1
    seed, site = 11, 6
    d = _lib.make_dropout(seed, site, 0.5)
    mask = torch.tensor(helpers.keep_mask(seed, site, 0.5, M, U)).double() * 2.0
    # ---- float64 reference with autograd
    Z64 = Z0.double().requires_grad_(True)
    U64 = Uw.double()
    hs, cs = torch.zeros(S, U, dtype=torch.float64), torch.zeros(S, U, dtype=torch.float64)
    loss = 0.0
    for t in range(steps):
        z = (Z64[rows[:, t]] + hs @ U64).view(S, U, 4)
        i, f, o = [(0.2 * z[..., k] + 0.5).clamp(0, 1) for k in (0, 1, 3)]
        cs = f * cs + i * torch.tanh(z[..., 2])
        hs = o * torch.tanh(cs)
        loss = loss + (hs * mask[rows[:, t]] * dY[rows[:, t], :U].double()).sum()
    loss.backward()
    dZref = Z64.grad
    # ---- device: fp32 forward scan for the saved gates / cell states, tensor-core reverse scan
    Z, Ud = Z0.cuda(), Uw.cuda()
    h, c = torch.empty(M, U, device="cuda"), torch.empty(M, U, device="cuda")
    _lib.check(lib.dj_lstm_scan_fwd(P(Z), P(h), P(c), None, P(Ud), S, steps, U, *m, 1, None))
    G16 = _gates16(Z)
    dZt = torch.zeros(M, 4 * U, device="cuda").bfloat16()
    dbt = torch.zeros(4 * U, device="cuda")
    dYd = dY.cuda()
    _lib.check(lib.dj_lstm_scan_tc_bwd(P(G16), P(c), P(dYd), ld, d, P(Uw.bfloat16().cuda()), P(dZt), P(dbt), S, steps, U,
                                       *m, 1, None))
    torch.cuda.synchronize()
    a, b = dZt.float().cpu().double(), dZref
    scale = float(b.abs().max())
    print(f"scan_tc_bwd vs float64 autograd [{axis}]: max |d| / max {float((a - b).abs().max()) / scale:.2e}, "
          f"mean |d| / mean {float((a - b).abs().mean() / b.abs().mean()):.2e}")
    assert float((a - b).abs().max()) / scale < 0.02 and float((a - b).abs().mean() / b.abs().mean()) < 0.01
    assert helpers.rel_err(dbt.cpu().numpy(), b.sum(0).numpy()) < 0.01


@pytest.mark.parametrize("Uprev,chosen", [(256, False), (256, True), (128, False), (512, True)])
@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("drop", [True, False])
def test_layer_input_fast_path_matches_numpy_twin(lib, Uprev, chosen, dtype, drop):
    """dj_layer_input (model.py:85,101-106,113-117) in the shapes the engine launches -- 16-bit hi + lo operands,
    byte-mode dropout on both sites or none -- against a NumPy restatement that replays the same masks: the hi
    operand must be the correctly rounded value bit for bit, the lo operand the rounded residual; the fp32 output of
    the generic kernel must equal the NumPy value exactly."""
    from music_generator_b200 import _lib
    B, T, N = 3, 5, 48
    rows, F = B * T * N, Uprev + (3 if chosen else 0)
    ld = (F + 31) // 32 * 32
    g = torch.Generator().manual_seed(Uprev + chosen)
    h = torch.randn(rows, Uprev, generator=g)
    sp = torch.tanh(torch.randn(B * T, F, generator=g))
    ch = (torch.rand(B, T, N, 3, generator=g) < 0.3).float()
    seed = 77
    dh = _lib.make_dropout(seed, 8, 0.5) if drop else _lib.NO_DROPOUT
    ds = _lib.make_dropout(seed, 9, 0.5) if drop else _lib.NO_DROPOUT
    dc = _lib.make_dropout(seed, 3, 0.2) if drop else _lib.NO_DROPOUT
    one = np.float32(1)
    mh = helpers.keep_mask(seed, 8, 0.5, rows, Uprev).astype(np.float32) * np.float32(2) if drop else one
    ms = helpers.keep_mask(seed, 9, 0.5, rows, F).astype(np.float32) * np.float32(2) if drop else one
    mc = helpers.keep_mask(seed, 3, 0.2, rows, 3).astype(np.float32) * np.float32(1.25) if drop else one
    ref = np.zeros((rows, ld), np.float32)
    ref[:, :Uprev] = h.numpy() * mh
    if chosen:
        c = (ch.numpy().reshape(rows, 3) * mc).astype(np.float32)
        sh = np.zeros_like(c)
        sh[1:] = c[:-1]
        sh[np.arange(rows) % N == 0] = 0          # note 0 has no previous note
        ref[:, Uprev:F] = sh
    ref[:, :F] += np.repeat(sp.numpy(), N, axis=0) * ms
    hd, spd, chd = h.cuda(), sp.cuda(), ch.cuda()
    out32 = torch.full((rows, ld), 7.0, device="cuda")
    args = lambda A, Alo, dt: (P(hd), Uprev, 0, T * N, dh, P(spd), F, ds, P(chd) if chosen else None, T * N * 3,
                               dc, B, T, P(A), None if Alo is None else P(Alo), ld, dt, None)
    _lib.check(lib.dj_layer_input(*args(out32, None, 0)))
    tdt, code = (torch.bfloat16, 1) if dtype == "bf16" else (torch.float16, 2)
    hi = torch.full((rows, ld), 7.0, device="cuda", dtype=tdt)
    lo = torch.full((rows, ld), 7.0, device="cuda", dtype=tdt)
    _lib.check(lib.dj_layer_input(*args(hi, lo, code)))
    torch.cuda.synchronize()
    assert np.array_equal(out32.cpu().numpy(), ref)
    r = torch.from_numpy(ref)
    want_hi = r.to(tdt)
    want_lo = (r - want_hi.float()).to(tdt)
    assert torch.equal(hi.cpu(), want_hi)
    assert torch.equal(lo.cpu(), want_lo)
