"""Phase timing of one train step without serialising the two backward streams: forward / backward / optimizer."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import music_generator_b200  # noqa
from music_generator_b200.config import ModelConfig
from music_generator_b200.engine import Engine
import dataset

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
e = Engine(ModelConfig(), precision="bf16"); e.init_params(0)
x, y = dataset.synthetic_all(B)
dev = [torch.tensor(a).cuda() for a in x] + [torch.tensor(y[0]).cuda()]
for i in range(3):
    e.train_step(*dev, seed=i)
torch.cuda.synchronize()
acc = np.zeros(5)
K = 10
for i in range(K):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record()
    ws = e.workspace(B, 128, e.precision, True)
    d = e._drops(True, 10 + i)
    e._last = dict(ws=ws, d=d, bf16=True, notes=dev[0], style=dev[3], B=B, T=128)
    e.forward_time(ws, dev[0], 128 * 48 * 3, dev[2], 128 * 16, B, 128, d, True, True, style=dev[3],
                   style_bstride=128 * 23, style_tstride=23)
    ev[1].record()
    e.forward_note(ws, ws.h[1], 0, 128 * 48, dev[1], 128 * 48 * 3, B, 128, d, True, True, dev[4])
    ev[2].record()
    e.backward()
    ev[3].record()
    e.nadam_step(1.0)
    ev[4].record()
    torch.cuda.synchronize()
    acc[:4] += [ev[j].elapsed_time(ev[j + 1]) for j in range(4)]
    acc[4] += ev[0].elapsed_time(ev[4])
print("forward time-axis %.3f  forward note-axis+heads %.3f  backward %.3f  nadam %.3f  total %.3f ms" % tuple(acc / K))

# ---- in-situ timeline of the backward pass (two streams kept): start offset and duration of every launch
names = set()
for L in ("note1", "note0", "time1", "time0"):
    for k in ("dj_lstm_scan_tc_bwd", "dj_gate_gemm_bf16", "dj_wgrad_gemm_bf16", "dj_style_bwd_reduce", "dj_gemm_simt", "dj_colsum"):
        names.add(f"{k}:bwd:{L}")
names |= {"dj_conv_bwd", "dj_head_finalize", "dj_gemm_simt", "dj_colsum", "dj_nadam_step"}
e.forward(*dev[:4], target=dev[4], train=True, seed=77)
torch.cuda.synchronize()
e.profile, e.profile_only = [], names
t0 = torch.cuda.Event(enable_timing=True); t0.record()
e.backward(); e.nadam_step(1.0)
t1 = torch.cuda.Event(enable_timing=True); t1.record()
torch.cuda.synchronize()
print("backward+nadam %.3f ms; launches (start offset, duration):" % t0.elapsed_time(t1))
for name, a, b in e.profile:
    print("  %-34s start %7.3f  dur %6.3f" % (name, t0.elapsed_time(a), a.elapsed_time(b)))
e.profile, e.profile_only = None, None
