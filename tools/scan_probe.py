"""Runs only the LSTM scan kernels at training shapes (for ncu / timing): bf16 single-pass and half hi+lo modes."""
import ctypes as C
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import music_generator_b200  # noqa
from music_generator_b200 import _lib

if os.environ.get("DJ_PROBE_LIB"):
    _lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), os.environ["DJ_PROBE_LIB"])
lib = _lib.load()
P = lambda t: C.c_void_p(t.data_ptr())
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
which = sys.argv[2] if len(sys.argv) > 2 else "all"
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["bf16", "half_split"]
T = 128
g = torch.Generator().manual_seed(0)
for axis, U in (("time", 256), ("note", 128)):
    M = B * T * 48
    Z0 = torch.randn(M, 4 * U, generator=g).cuda()
    Uw = (torch.randn(U, 4 * U, generator=g) * 0.06).cuda()
    S, steps, m = (B * 48, T, (48, T * 48, 1, 48)) if axis == "time" else (B * T, 48, (1, 48, 0, 1))
    h, c = torch.empty(M, U, device="cuda"), torch.empty(M, U, device="cuda")
    dY = torch.randn(M, U, device="cuda") * 0.01
    dZ = torch.empty(M, 4 * U, device="cuda").bfloat16()
    db = torch.zeros(4 * U, device="cuda")
    G16 = torch.rand(M, 4 * U, device="cuda").half()
    for mode in modes:
        dt = torch.bfloat16 if mode == "bf16" else torch.float16
        fmt = 1 if mode == "bf16" else 2
        Ut = Uw.t().contiguous().to(dt)
        Ut_lo = (Uw.t().contiguous() - Ut.float()).to(dt)
        Un = Uw.bfloat16()
        hp = torch.zeros(M, U, device="cuda").to(dt)
        for rep in range(3):
            Z = Z0.clone()
            e0, e1, e2, e3 = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            e0.record()
            if which in ("all", "fwd"):
                _lib.check(lib.dj_lstm_scan_tc_fwd(P(Z), P(G16), P(h), P(c), P(hp), P(Ut), P(Ut_lo) if mode != "bf16" else None, fmt,
                                                   S, steps, U, *m, 1, None))
            e1.record()
            if which in ("all", "bwd"):
                _lib.check(lib.dj_lstm_scan_tc_bwd(P(G16), P(c), P(dY), U, _lib.NO_DROPOUT, P(Un), P(dZ), P(db), S, steps, U,
                                                   *m, 1, None))
            e2.record()
            torch.cuda.synchronize()
        print(f"{axis} [{mode}]: B={B} tc_fwd {e0.elapsed_time(e1):.3f} ms ({1e3 * e0.elapsed_time(e1) / steps:.2f} us/step)  "
              f"tc_bwd {e1.elapsed_time(e2):.3f} ms ({1e3 * e1.elapsed_time(e2) / steps:.2f} us/step)", flush=True)
