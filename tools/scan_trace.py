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
    tr = trace.cpu().numpy().reshape(2, 512, 16)
    for slot in range(2):
        a = tr[slot, :steps].astype(np.float64)
        order = range(steps - 2, 1, -1) if reverse else range(2, steps - 1)
        rows = []
        prev = None
        for t in order:
            if prev is not None:
                # one step = from accumulator-ready of step `prev` to accumulator-ready of step t
                e8p, e9, e10, e11 = a[prev, 8], a[prev, 9], a[prev, 10], a[prev, 11]
                w2, w3 = a[prev, 2], a[prev, 3]
                w0, w1, e8 = a[t, 0], a[t, 1], a[t, 8]
                rows.append([e9 - e8p, e10 - e9, e11 - e10, w2 - e10, w3 - w2, w0 - w3, w1 - w0, e8 - w1, e8 - e8p])
            prev = t
        if not reverse:
            ex = a[4:steps - 4, 12:16]
            print(f"   fwd epilogue split (sum over chunks): tmem_ld {np.median(ex[:, 0]):.0f}  z landed {np.median(ex[:, 1]):.0f}"
                  f"  load issue {np.median(ex[:, 2]):.0f}  math+stores {np.median(ex[:, 3]):.0f}")
        else:
            sel = a[4:steps - 4]
            print(f"   bwd epilogue split: tmem_ld {np.median(sel[:, 12] - sel[:, 8]):.0f}  math+stores {np.median(sel[:, 9] - sel[:, 12]):.0f}"
                  f"  fences {np.median(sel[:, 13] - sel[:, 9]):.0f}  arrive {np.median(sel[:, 10] - sel[:, 13]):.0f}"
                  f"  issue_loads+cl_wait {np.median(sel[:, 11] - sel[:, 10]):.0f}")
        r = np.array(rows[4:-4])
        lab = ["epilogue", "fence", "cl_barrier(epi)", "cl_barrier(issuer, from fence)", "tma issue", "tma flight",
               "mma issue", "mma->acc ready", "STEP"]
        print(f"{name} slot {slot}: " + "  ".join(f"{l} {m:.0f}" for l, m in zip(lab, np.median(r, axis=0))) + " [cycles]")


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
    for rep in range(2):
        Z = Z0.clone()
        trace.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.dj_lstm_scan_tc_fwd(P(Z), P(h), P(c), P(hp), P(Ut), S, steps, U, *m, 1, None))
        e1.record()
        torch.cuda.synchronize()
        if rep == 1:
            print(f"{axis} fwd B={B}: {e0.elapsed_time(e1):.3f} ms  {1e3 * e0.elapsed_time(e1) / steps:.2f} us/step")
            report(f"{axis} fwd", steps, False)
        trace.zero_()
        e0.record()
        _lib.check(lib.dj_lstm_scan_tc_bwd(P(Z), P(c), P(dY), U, _lib.NO_DROPOUT, P(Un), P(dZ), P(db), S, steps, U, *m, 1, None))
        e1.record()
        torch.cuda.synchronize()
        if rep == 1:
            print(f"{axis} bwd B={B}: {e0.elapsed_time(e1):.3f} ms  {1e3 * e0.elapsed_time(e1) / steps:.2f} us/step")
            report(f"{axis} bwd", steps, True)
