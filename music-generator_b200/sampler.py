"""On-device autoregressive generation (reference generate.py:13-121).

Per generated timestep the host only enqueues kernels: the full 128-step
time-axis window recompute from zero state (exact reference semantics,
generate.py:106-109) at fp32 precision, then ONE persistent sampler launch for
the 48 notes.  No host synchronisation happens inside the loop; events are read
back once at the end.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from ._lib import DJ_F32, NO_DROPOUT
from .engine import Engine, N, Workspace, _ptr, _stream

PREDICT_CHUNK = 32   # keras Model.predict default batch_size
REFERENCE_STREAM_MAX = 8   # sequences one sampler cluster walks in np.random call order (generate.py default: 3)


class DeviceGeneration:
    """State of G `MusicGeneration` objects (generate.py:17-30) held in HBM."""

    def __init__(self, eng: Engine, styles: Sequence[np.ndarray], num_steps: int, default_temp: float = 1.0):
        cfg = eng.cfg
        self.eng, self.G, self.steps, self.L = eng, len(styles), num_steps, cfg.seq_len
        G, L, dev = self.G, self.L, eng.dev
        f32 = dict(dtype=torch.float32, device=dev)
        # sliding windows as one history buffer: window(t) = hist[:, t:t+L]
        self.hist_notes = torch.zeros(G, L + num_steps, N, 3, **f32)          # generate.py:18
        beat = torch.zeros(G, L + num_steps, cfg.notes_per_bar)               # :19 zeros, then compute_beat(t)
        ts = torch.arange(num_steps)
        beat[:, L + ts, ts % cfg.notes_per_bar] = 1.0                        # generate.py:75
        self.hist_beat = beat.to(dev)
        self.style = torch.tensor(np.stack([np.asarray(s, dtype=np.float64) for s in styles]), **f32)   # :20
        self.temperature = torch.full((G,), float(default_temp), dtype=torch.float64, device=dev)
        self.silent_time = torch.full((G,), cfg.notes_per_bar, dtype=torch.int32, device=dev)     # :24
        self.default_temp = float(default_temp)
        self.events = torch.zeros(G, N, 3, **f32)
        self.margin = torch.full((G,), 1e300, dtype=torch.float64, device=dev)
        self.cursor = torch.zeros(1, dtype=torch.int64, device=dev)
        # eng.gen_tc: the window recompute on the tensor cores at fp32 grade (half hi+lo operands); else CUDA-core fp32
        self.ws_time = Workspace(cfg, G, L, "gen" if eng.gen_tc else "fp32", False, dev)
        self.ws_note = Workspace(cfg, G, 1, "fp32", False, dev)
        self.zero_chosen = torch.zeros(G, 1, N, 3, **f32)
        self.probs = torch.zeros(num_steps, G, N, 3, **f32)
        self.temp_trace = torch.zeros(num_steps, G, dtype=torch.float64, device=dev)   # temperature used at each step
        # style is constant in time: embed/project once (model.py:141-142, 77-79, 113-115)
        eng._style(self.ws_time, self.style, cfg.num_styles, 0, G, L)
        eng._style(self.ws_note, self.style, cfg.num_styles, 0, G, 1)

    def step(self, t: int, uniforms: torch.Tensor, stream_mode: int, forced: Optional[torch.Tensor] = None):
        eng, cfg, G, L = self.eng, self.cfg, self.G, self.L
        d = {s: NO_DROPOUT for s in range(1, 13)}
        bn = (L + self.steps) * N * 3
        bb = (L + self.steps) * cfg.notes_per_bar
        notes_w = self.hist_notes.view(-1)[t * N * 3:]
        beat_w = self.hist_beat.view(-1)[t * cfg.notes_per_bar:]
        eng.forward_time(self.ws_time, notes_w, bn, beat_w, bb, G, L, d, False, False, style_done=True)
        ws = self.ws_note
        L2, P = eng.layers[2], eng.params
        un = cfg.note_axis_units
        # note layer 0 projection of [time_out(last step) + style, style] for all 48 notes
        eng._call("dj_layer_input", _ptr(self.ws_time.h[1]), cfg.time_axis_units, (L - 1) * N, L * N, NO_DROPOUT,
                  _ptr(ws.sp[2]), L2["F"], NO_DROPOUT, _ptr(self.zero_chosen), N * 3, NO_DROPOUT, G, 1,
                  _ptr(ws.A[2]), None, ws.ld[2], DJ_F32, _stream())
        eng._gate_gemm(2, ws, False, G * N)
        W0 = P["note0.lstm.W"]
        W0c = W0[cfg.time_axis_units:]            # rows of the 3 chosen channels (contiguous tail)
        u = uniforms if stream_mode == 0 else uniforms[t]
        self.temp_trace[t].copy_(self.temperature)
        eng._call("dj_gen_sample", _ptr(ws.Z[2]), _ptr(W0c), _ptr(P["note0.lstm.U"]), _ptr(P["note1.lstm.W"]),
                  _ptr(P["note1.lstm.U"]), _ptr(P["note1.lstm.b"]), _ptr(ws.sp[3]), _ptr(P["note_dense.W"]),
                  _ptr(P["note_dense.b"]), _ptr(P["volume_dense.W"]), _ptr(P["volume_dense.b"]), un, G, _ptr(u),
                  _ptr(self.cursor), stream_mode, _ptr(self.temperature), _ptr(self.silent_time),
                  self.default_temp, eng.hard, _ptr(self.events), _ptr(self.probs[t]), _ptr(self.margin), _stream())
        # generate.py:73: push next_note into the window (device-to-device copy, no sync)
        self.hist_notes[:, L + t].copy_(self.events if forced is None else forced)

    @property
    def cfg(self):
        return self.eng.cfg

    def results(self) -> np.ndarray:
        """[steps, G, 48, 3] events, as generate.py:121 yields them."""
        return self.hist_notes[:, self.L:].permute(1, 0, 2, 3).contiguous().cpu().numpy()


class GenerationRun:
    """A whole generation job on one GPU: the sequences in predict-chunks of 32 (the chunk scopes the pitch_bins
    scramble, model.py:43-49, exactly as Keras' predict(batch_size=32) does), the uniform stream on the device, and a
    `step(t)` that enqueues one generated timestep for every chunk without synchronising."""

    def __init__(self, eng: Engine, styles: Sequence[np.ndarray], num_steps: int, uniforms: np.ndarray,
                 stream_mode: int = 0, default_temp: float = 1.0, forced_events: Optional[np.ndarray] = None):
        G = len(styles)
        u = torch.tensor(np.ascontiguousarray(uniforms, dtype=np.float64), device=eng.dev)
        if stream_mode == 1:
            assert u.shape == (num_steps, G, N, 2), u.shape
            self.chunks = [(s, min(s + PREDICT_CHUNK, G)) for s in range(0, G, PREDICT_CHUNK)]
        else:
            if G > REFERENCE_STREAM_MAX:
                raise ValueError("reference stream order supports at most %d sequences" % REFERENCE_STREAM_MAX)
            self.chunks = [(0, G)]
        self.stream_mode, self.steps = stream_mode, num_steps
        self.gens = [DeviceGeneration(eng, styles[a:b], num_steps, default_temp) for a, b in self.chunks]
        self.us = [u if stream_mode == 0 else u[:, a:b].contiguous() for a, b in self.chunks]
        self.forced = None if forced_events is None else torch.tensor(forced_events, dtype=torch.float32, device=eng.dev)

    def step(self, t: int) -> None:
        for gen, uc, (a, b) in zip(self.gens, self.us, self.chunks):
            gen.step(t, uc, self.stream_mode, None if self.forced is None else self.forced[t, a:b])

    def finish(self):
        gens = self.gens
        info = dict(probs=np.concatenate([g.probs.cpu().numpy() for g in gens], axis=1),
                    min_margin=min(float(g.margin.min().item()) for g in gens),
                    uniforms_used=int(gens[0].cursor.item()),
                    temperature_trace=np.concatenate([g.temp_trace.cpu().numpy() for g in gens], axis=1),
                    temperature=np.concatenate([g.temperature.cpu().numpy() for g in gens]),
                    silent_time=np.concatenate([g.silent_time.cpu().numpy() for g in gens]))
        return np.concatenate([g.results() for g in gens], axis=1), info


def generate_events(eng: Engine, styles: Sequence[np.ndarray], num_steps: int, uniforms: np.ndarray,
                    stream_mode: int = 0, default_temp: float = 1.0, forced_events: Optional[np.ndarray] = None):
    """Run the whole generation on device.

    stream_mode 0: `uniforms` is the flat float64 stream consumed in reference
    order (generate.py:112-118; at most 8 sequences, one GPU);  1: indexed
    stream uniforms[t, g, n, 2] (shardable across GPUs).
    forced_events [steps, G, 48, 3]: lock-step mode for parity tests -- decisions
    are made and recorded but the forced event is what enters the window.
    Returns (events [steps,G,48,3], info dict with probs / min_margin / uniforms_used / temperature trace).
    """
    run = GenerationRun(eng, styles, num_steps, uniforms, stream_mode, default_temp, forced_events)
    for t in range(num_steps):
        run.step(t)
    return run.finish()
