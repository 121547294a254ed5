"""Per-step timeline of the tensor-core scan kernels (debug build with -DDJ_TRACE).

  DJ_OUT=../libdeepj_trace.so DJ_BUILD_DIR=../build_trace DJ_NVCC_EXTRA=-DDJ_TRACE bash music-generator_b200/csrc/build.sh
  python tools/scan_trace.py 64

Stamps (clock64 of the traced CTA's SM), per step t:
  issuer lane:   0 operand landed (mbarrier)   1 MMAs issued + commit   2 cluster barrier passed   3 TMA multicast issued
  epilogue lane: 8 accumulator ready   9 epilogue math + stores issued   10 after fence.proxy.async   11 cluster barrier passed
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import music_generator_b200  # noqa
from music_generator_b200 import _lib

_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), os.environ.get("DJ_TRACE_LIB", "libdeepj_trace.so"))
lib = _lib.load()
lib.dj_debug_trace_set.restype = C.c_int
lib.dj_debug_trace_set.argtypes = [C.c_void_p]
P = lambda t: C.c_void_p(t.data_ptr())
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = 128
g = torch.Generator().manual_seed(0)
trace = torch.zeros(2 * 512 * 16, dtype=torch.int64, device="cuda")
assert lib.dj_debug_trace_set(P(trace)) == 0


def report(name, steps, reverse):
    """Median of every stamp relative to stamp 8 (first accumulator ready) of the same step, and the step period."""
    tr = trace.cpu().numpy().reshape(2, 512, 16)
    for slot in range(2):
        a = tr[slot, :steps].astype(np.float64)
        sel = a[6:steps - 6]
        nxt = a[5:steps - 7] if reverse else a[7:steps - 5]
        period = np.median(nxt[:, 8] - sel[:, 8])
        rel = {k: np.median(sel[:, k] - sel[:, 8]) for k in range(16) if k != 8 and np.all(sel[:, k] != 0)}
        print(f"{name} slot {slot}: STEP {period:.0f} | " + "  ".join(f"k{k}:{v:+.0f}" for k, v in sorted(rel.items(), key=lambda kv: kv[1])))


AXES = [a for a in (("time", 256), ("note", 128)) if len(sys.argv) < 3 or sys.argv[2] == a[0]]
for axis, U in AXES:
    M = B * T * 48
    Z0 = torch.randn(M, 4 * U, generator=g).cuda()
    Uw = (torch.randn(U, 4 * U, generator=g) * 0.06).cuda()
    Ut = Uw.t().contiguous().bfloat16()
    Un = Uw.bfloat16()
    S, steps, m = (B * 48, T, (48, T * 48, 1, 48)) if axis == "time" else (B * T, 48, (1, 48, 0, 1))
    h, c = torch.empty(M, U, device="cuda"), torch.empty(M, U, device="cuda")
    hp = torch.zeros(M, U, device="cuda").bfloat16()
    dY = torch.randn(M, U, device="cuda") * 0.01
    dZ = torch.empty(M, 4 * U, device="cuda").bfloat16()
    db = torch.zeros(4 * U, device="cuda")
    G16 = torch.rand(M, 4 * U, device="cuda").half()
    for rep in range(2):
        Z = Z0.clone()
        trace.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.dj_lstm_scan_tc_fwd(P(Z), P(G16), P(h), P(c), P(hp), P(Ut), None, 1, S, steps, U, *m, 1, None))
        e1.record()
        torch.cuda.synchronize()
        if rep == 1:
            print(f"{axis} fwd B={B}: {e0.elapsed_time(e1):.3f} ms  {1e3 * e0.elapsed_time(e1) / steps:.2f} us/step")
            report(f"{axis} fwd", steps, False)
        trace.zero_()
        e0.record()
        _lib.check(lib.dj_lstm_scan_tc_bwd(P(G16), P(c), P(dY), U, _lib.NO_DROPOUT, P(Un), P(dZ), P(db), S, steps, U, *m, 1, None))
        e1.record()
        torch.cuda.synchronize()
        if rep == 1:
            print(f"{axis} bwd B={B}: {e0.elapsed_time(e1):.3f} ms  {1e3 * e0.elapsed_time(e1) / steps:.2f} us/step")
            report(f"{axis} bwd", steps, True)
