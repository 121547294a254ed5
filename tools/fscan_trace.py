"""Per-step timeline of the fp32 forward scan (debug build with -DDJ_TRACE): thread 0 of block 0.
stamps: 0 step start, 1 h_{t-1} landed (mbarrier), 2 k-loop done, 3 activations + stores done (before the sends)"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import music_generator_b200  # noqa
from music_generator_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), "libdeepj_trace.so")
lib = _lib.load()
lib.dj_debug_ftrace_set.argtypes = [C.c_void_p]
P = lambda t: C.c_void_p(t.data_ptr())
trace = torch.zeros(512 * 8, dtype=torch.int64, device="cuda")
assert lib.dj_debug_ftrace_set(P(trace)) == 0
U, T = 256, 128
for G in (1, 32):
    S = 48 * G
    M = G * T * 48
    Z = torch.randn(M, 4 * U).cuda()
    Uw = (torch.randn(U, 4 * U) * 0.06).cuda()
    h = torch.empty(M, U, device="cuda")
    for rep in range(2):
        trace.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.dj_lstm_scan_fwd(P(Z), P(h), None, None, P(Uw), S, T, U, 48, T * 48, 1, 48, 1, None))
        e1.record(); torch.cuda.synchronize()
    a = trace.cpu().numpy().reshape(512, 8).astype(np.float64)[8:T - 8]
    nxt = trace.cpu().numpy().reshape(512, 8).astype(np.float64)[9:T - 7]
    print(f"G={G} S={S}: {e0.elapsed_time(e1):.3f} ms | per step (cycles): wait h {np.median(a[:,1]-a[:,0]):.0f}  k-loop {np.median(a[:,2]-a[:,1]):.0f}"
          f"  activations+stores {np.median(a[:,3]-a[:,2]):.0f}  sends+loop {np.median(nxt[:,0]-a[:,3]):.0f}  STEP {np.median(nxt[:,0]-a[:,0]):.0f}")
