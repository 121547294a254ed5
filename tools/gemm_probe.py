"""Times the tcgen05 GEMMs at the training shapes (B=64)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import music_generator_b200  # noqa
from music_generator_b200 import _lib
lib = _lib.load(); P = lambda t: C.c_void_p(t.data_ptr())
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
M = B * 128 * 48
g = torch.Generator().manual_seed(0)
def tm(fn, n=5):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, K, N in (("time0 fwd", 96, 1024), ("time1 fwd", 256, 1024), ("note0 fwd", 288, 512), ("note1 fwd", 128, 512),
                   ("time1 dgrad", 1024, 256), ("note0 dgrad", 512, 259)):
    A = torch.randn(M, K, generator=g).bfloat16().cuda(); Bt = torch.randn(N, K, generator=g).bfloat16().cuda()
    ldc = (N + 31) // 32 * 32
    Cc = torch.empty(M, ldc, device="cuda"); bias = torch.zeros(ldc, device="cuda")
    ms = tm(lambda: _lib.check(lib.dj_gate_gemm_bf16(P(A), K, P(Bt), K, P(Cc), ldc, P(bias) if N % 4 == 0 else None, M, N, K, None)))
    gb = (M * K * 2 + M * N * 4) / 1e9
    print(f"{name:12s} M={M} N={N} K={K}: {ms:.3f} ms  {gb / ms:.2f} TB/s  {2 * M * N * K / ms / 1e9:.0f} TFLOP/s", flush=True)
for name, Ka, Nb in (("time1 dW", 256, 1024), ("note0 dW", 259, 512)):
    A = torch.randn(M, (Ka + 31) // 32 * 32, generator=g).bfloat16().cuda(); Bm = torch.randn(M, Nb, generator=g).bfloat16().cuda()
    Cc = torch.zeros(Ka, Nb, device="cuda")
    ms = tm(lambda: _lib.check(lib.dj_wgrad_gemm_bf16(P(A), A.shape[1], P(Bm), Nb, P(Cc), Nb, Ka, Nb, M, None)))
    print(f"{name:12s} M={M} Ka={Ka} Nb={Nb}: {ms:.3f} ms  {(M * (Ka + Nb) * 2) / 1e9 / ms:.2f} TB/s  {2 * M * Ka * Nb / ms / 1e9:.0f} TFLOP/s", flush=True)
