"""CPU oracle for the DeepJ biaxial-LSTM hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain PyTorch-CPU tensor code, the arithmetic that the
reference expresses as a Keras-2 / TensorFlow-1 graph.  It is imported only by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` -- never by the product path (the product
path is the CUDA library in ``music-generator_b200/csrc`` and raises if that
library is missing).

PARITY UNPINNED: the reference's own tests (``test.py``) only cover the MIDI
codec; it ships no golden vector, fixture or known-answer test for
``model.py`` / ``generate.py``; Keras/TensorFlow (un-vendored, un-pinned in
``requirements.txt:1-2``; era Keras 2.0-2.1 / TF <= 1.4 by ``scripts/cuda.sh:9``)
cannot be installed here.  The oracle therefore restates the *published*
Keras-2 semantics at the reference's call sites; the version-sensitive ones
are explicit switches (``recurrent_activation``, ``nadam_eps``).  The pins we
create ourselves are the hand-computed quirk tests in ``tests/test_oracle.py``
and the literal-loop second restatement in ``oracle/literal.py``.

Reference call sites followed (all paths relative to /root/reference):
  model.py:14-20    primary_loss
  model.py:22-49    pitch_pos_in_f / pitch_class_in_f / pitch_bins_f
  model.py:51-89    time_axis
  model.py:91-126   note_axis
  model.py:128-169  build_models (input dropout, style Dense, 3 models)
  generate.py:13-121 MusicGeneration / apply_temperature / generate
  dataset.py:14-26  compute_beat / compute_genre
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch


# --------------------------------------------------------------------------
# configuration (constants.py:42-77)
# --------------------------------------------------------------------------
@dataclass(frozen=True)
class Config:
    num_styles: int = 23          # constants.py:42
    num_octaves: int = 4          # constants.py:50
    octave: int = 12              # constants.py:51
    notes_per_bar: int = 16       # constants.py:63
    seq_len: int = 128            # constants.py:67
    octave_units: int = 64        # constants.py:70
    style_units: int = 64         # constants.py:71
    note_units: int = 3           # constants.py:72
    time_axis_units: int = 256    # constants.py:73
    note_axis_units: int = 128    # constants.py:74
    time_axis_layers: int = 2     # constants.py:76
    note_axis_layers: int = 2     # constants.py:77

    @property
    def num_notes(self) -> int:   # constants.py:54-56
        return self.num_octaves * self.octave

    @property
    def feat0(self) -> int:       # model.py:61-67: pos 1 + class 12 + bins 1 + conv + beat
        return 1 + self.octave + 1 + self.octave_units + self.notes_per_bar

    def time_in_dims(self) -> List[int]:
        return [self.feat0] + [self.time_axis_units] * (self.time_axis_layers - 1)

    def note_in_dims(self) -> List[int]:
        return [self.time_axis_units + self.note_units] + \
            [self.note_axis_units] * (self.note_axis_layers - 1)


def param_shapes(cfg: Config) -> "Dict[str, tuple]":
    """The 28 weight tensors in Keras layouts, in model.py creation order."""
    s: Dict[str, tuple] = {}
    s["style.W"] = (cfg.num_styles, cfg.style_units)            # model.py:141
    s["style.b"] = (cfg.style_units,)
    s["conv.W"] = (2 * cfg.octave, cfg.note_units, cfg.octave_units)   # model.py:56
    s["conv.b"] = (cfg.octave_units,)
    for l, f in enumerate(cfg.time_in_dims()):                   # model.py:75-85
        u = cfg.time_axis_units
        s[f"time{l}.sd.W"] = (cfg.style_units, f)
        s[f"time{l}.sd.b"] = (f,)
        s[f"time{l}.lstm.W"] = (f, 4 * u)
        s[f"time{l}.lstm.U"] = (u, 4 * u)
        s[f"time{l}.lstm.b"] = (4 * u,)
    for l, f in enumerate(cfg.note_in_dims()):                   # model.py:108-123
        u = cfg.note_axis_units
        s[f"note{l}.sd.W"] = (cfg.style_units, f)
        s[f"note{l}.sd.b"] = (f,)
        s[f"note{l}.lstm.W"] = (f, 4 * u)
        s[f"note{l}.lstm.U"] = (u, 4 * u)
        s[f"note{l}.lstm.b"] = (4 * u,)
    s["note_dense.W"] = (cfg.note_axis_units, 2)                # model.py:94
    s["note_dense.b"] = (2,)
    s["volume_dense.W"] = (cfg.note_axis_units, 1)              # model.py:95
    s["volume_dense.b"] = (1,)
    return s


def init_params(cfg: Config, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Keras default initialisers: glorot_uniform kernels, orthogonal recurrent
    kernels, zero biases with the LSTM forget slice set to 1 (unit_forget_bias)."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shp in param_shapes(cfg).items():
        if name.endswith(".b"):
            t = torch.zeros(shp, dtype=torch.float64)
            if ".lstm." in name:
                u = shp[0] // 4
                t[u:2 * u] = 1.0
        elif name.endswith("lstm.U"):
            u = shp[0]
            blocks = []
            for _ in range(4):   # one orthogonal [u,u] block per gate is a valid orthogonal-rows init
                a = torch.randn(u, u, generator=g, dtype=torch.float64)
                q, r = torch.linalg.qr(a)
                q = q * torch.sign(torch.diagonal(r))
                blocks.append(q)
            t = torch.cat(blocks, dim=1) * 0.5   # keep 4-gate concat row-norm 1
        else:
            if len(shp) == 3:   # conv: fan_in = k*in, fan_out = k*out
                fan_in, fan_out = shp[0] * shp[1], shp[0] * shp[2]
            else:
                fan_in, fan_out = shp
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            t = (torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * lim
        out[name] = t.to(dtype)
    return out


# --------------------------------------------------------------------------
# layer primitives (Keras semantics)
# --------------------------------------------------------------------------
def hard_sigmoid(x: torch.Tensor) -> torch.Tensor:
    # Keras-2 backend hard_sigmoid: clip(0.2*x + 0.5, 0, 1)
    return torch.clamp(0.2 * x + 0.5, 0.0, 1.0)


def lstm_seq(x: torch.Tensor, W, U, b, recurrent_activation: str = "hard_sigmoid",
             return_state: bool = False):
    """keras.layers.LSTM(units, return_sequences=True) on x [S, steps, F], zero
    initial state, gate order i,f,c,o (model.py:84,120)."""
    S, steps, _ = x.shape
    u = U.shape[0]
    act = hard_sigmoid if recurrent_activation == "hard_sigmoid" else torch.sigmoid
    h = x.new_zeros(S, u)
    c = x.new_zeros(S, u)
    zxs = (x @ W + b).unbind(1)
    hs = []
    for t in range(steps):
        z = zxs[t] + h @ U
        i = act(z[:, 0 * u:1 * u])
        f = act(z[:, 1 * u:2 * u])
        g = torch.tanh(z[:, 2 * u:3 * u])
        o = act(z[:, 3 * u:4 * u])
        c = f * c + i * g
        h = o * torch.tanh(c)
        hs.append(h)
    out = torch.stack(hs, dim=1)
    if return_state:
        return out, (h, c)
    return out


def dropout(x: torch.Tensor, mask: Optional[torch.Tensor], rate: float) -> torch.Tensor:
    """tf.nn.dropout: x * mask / keep_prob.  mask None = inference (identity)."""
    if mask is None:
        return x
    return x * mask.to(x.dtype) / (1.0 - rate)


def conv1d_same(x: torch.Tensor, W: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """TimeDistributed(Conv1D(O, k, padding='same')) over the note axis
    (model.py:56).  TF 'SAME' with even k pads (k-1)//2 left, k//2 right; it is
    a cross-correlation (no kernel flip).  x [R, N, C], W [k, C, O]."""
    k = W.shape[0]
    left, right = (k - 1) // 2, k // 2
    xp = torch.nn.functional.pad(x, (0, 0, left, right))
    cols = xp.unfold(1, k, 1)                       # [R, N, C, k]
    return torch.einsum("rnck,kco->rno", cols, W) + b


def pitch_bins(x: torch.Tensor, cfg: Config) -> torch.Tensor:
    """model.py:43-49 verbatim in torch: stack 12 strided slices on a NEW
    leading axis, sum the octave axis, tile 4x on the leading axis, then a raw
    reshape to [B, T, 48, 1] (the quirk: no transpose)."""
    B, T = x.shape[0], x.shape[1]
    stacked = torch.stack([x[:, :, i::cfg.octave, 0] for i in range(cfg.octave)], dim=0)
    bins = stacked.sum(dim=3)                          # [12, B, T]
    bins = bins.repeat(cfg.num_octaves, 1, 1)          # tf.tile -> [48, B, T]
    return bins.reshape(B, T, cfg.num_notes, 1)


def primary_loss(y_true: torch.Tensor, y_pred: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """model.py:14-20 with the Keras TF-backend losses: clip to [eps,1-eps], go
    to logits, sigmoid-CE; per-note-axis mean; Keras then means the [B,T] map."""
    def bce(t, o):
        o = torch.clamp(o, eps, 1.0 - eps)
        x = torch.log(o / (1.0 - o))
        return (torch.clamp(x, min=0) - x * t + torch.log1p(torch.exp(-x.abs()))).mean(-1)
    played = y_true[..., 0]
    bce_note = bce(y_true[..., 0], y_pred[..., 0])
    bce_replay = bce(y_true[..., 1], played * y_pred[..., 1] + (1 - played) * y_true[..., 1])
    d = y_true[..., 2] - (played * y_pred[..., 2] + (1 - played) * y_true[..., 2])
    mse = (d * d).mean(-1)
    return (bce_note + bce_replay + mse).mean()


# --------------------------------------------------------------------------
# model.py forward
# --------------------------------------------------------------------------
DROPOUT_SITES = ("D1", "D2", "D3", "D4", "D5", "D6", "D7", "D8", "D9", "D10", "D11", "D12")


def dropout_site_shapes(cfg: Config, B: int, T: int) -> Dict[str, tuple]:
    N = cfg.num_notes
    ut, un = cfg.time_axis_units, cfg.note_axis_units
    tdims, ndims = cfg.time_in_dims(), cfg.note_in_dims()
    shp = {"D1": (B, T, N, cfg.note_units), "D2": (B, T, cfg.notes_per_bar),
           "D3": (B, T, N, cfg.note_units), "D4": (B, T, N, cfg.octave_units)}
    # all per-note masks are stored in [B,T,N,F] row order (the oracle permutes
    # as the reference does; the mask index is defined on the canonical order)
    shp["D5"] = (B, T, N, tdims[0]); shp["D6"] = (B, T, N, ut)
    shp["D7"] = (B, T, N, tdims[1]) if len(tdims) > 1 else None
    shp["D8"] = (B, T, N, ut) if len(tdims) > 1 else None
    shp["D9"] = (B, T, N, ndims[0]); shp["D10"] = (B, T, N, un)
    shp["D11"] = (B, T, N, ndims[1]) if len(ndims) > 1 else None
    shp["D12"] = (B, T, N, un) if len(ndims) > 1 else None
    return shp


def style_embed(p, style_in):
    return style_in @ p["style.W"] + p["style.b"]             # model.py:141-142 (linear)


def time_axis_forward(p, cfg: Config, notes, beat, style, masks=None, dropout_rate=0.5,
                      recurrent_activation="hard_sigmoid", taps: Optional[dict] = None):
    """model.py:51-89.  notes are post-input-dropout [B,T,N,3]; beat [B,T,16];
    style = style embedding [B,T,64].  Returns time_out [B,T,N,Ut]."""
    masks = masks or {}
    B, T, N = notes.shape[0], notes.shape[1], cfg.num_notes
    dt = notes.dtype
    conv = conv1d_same(notes.reshape(B * T, N, cfg.note_units), p["conv.W"], p["conv.b"])
    conv = torch.tanh(conv).reshape(B, T, N, cfg.octave_units)
    conv = dropout(conv, masks.get("D4"), dropout_rate)
    pos = (torch.arange(N, dtype=torch.float32) / N).to(dt).reshape(1, 1, N, 1).expand(B, T, N, 1)
    pcl = torch.zeros(N, cfg.octave, dtype=dt)
    pcl[torch.arange(N), torch.arange(N) % cfg.octave] = 1
    pcl = pcl.reshape(1, 1, N, cfg.octave).expand(B, T, N, cfg.octave)
    bins = pitch_bins(notes, cfg)
    beat_r = beat.reshape(B, T, 1, -1).expand(B, T, N, beat.shape[-1])
    x = torch.cat([pos, pcl, bins, conv, beat_r], dim=-1)      # [B,T,N,94]
    if taps is not None:
        taps["features"] = x
    x = x.permute(0, 2, 1, 3)                                  # [B,N,T,F]
    site = [("D5", "D6"), ("D7", "D8")]
    for l in range(cfg.time_axis_layers):
        sp = torch.tanh(style @ p[f"time{l}.sd.W"] + p[f"time{l}.sd.b"])   # [B,T,F]
        sp = sp.reshape(B, T, 1, -1).expand(B, T, N, sp.shape[-1])
        sp = dropout(sp, masks.get(site[l][0]) if l < 2 else None, dropout_rate)
        x = x + sp.permute(0, 2, 1, 3)
        if taps is not None:
            taps[f"time{l}.in"] = x.permute(0, 2, 1, 3)
        F = x.shape[-1]
        h = lstm_seq(x.reshape(B * N, T, F), p[f"time{l}.lstm.W"], p[f"time{l}.lstm.U"],
                     p[f"time{l}.lstm.b"], recurrent_activation)
        x = h.reshape(B, N, T, -1)
        if taps is not None:
            taps[f"time{l}.h"] = x.permute(0, 2, 1, 3)
        m = masks.get(site[l][1]) if l < 2 else None
        x = dropout(x, None if m is None else m.permute(0, 2, 1, 3), dropout_rate)
    return x.permute(0, 2, 1, 3)                               # [B,T,N,Ut]


def note_axis_forward(p, cfg: Config, time_out, chosen, style, masks=None, dropout_rate=0.5,
                      recurrent_activation="hard_sigmoid", taps: Optional[dict] = None):
    """model.py:91-126.  chosen is post-input-dropout [B,T,N,3] (3 channels are
    kept: the Reshape(-1) at model.py:104 is a no-op)."""
    masks = masks or {}
    B, T, N = time_out.shape[0], time_out.shape[1], cfg.num_notes
    shift = torch.nn.functional.pad(chosen[:, :, :-1, :], (0, 0, 1, 0))   # model.py:101
    x = torch.cat([time_out, shift], dim=3)
    site = [("D9", "D10"), ("D11", "D12")]
    for l in range(cfg.note_axis_layers):
        sp = torch.tanh(style @ p[f"note{l}.sd.W"] + p[f"note{l}.sd.b"])
        sp = sp.reshape(B, T, 1, -1).expand(B, T, N, sp.shape[-1])
        sp = dropout(sp, masks.get(site[l][0]) if l < 2 else None, dropout_rate)
        x = x + sp
        if taps is not None:
            taps[f"note{l}.in"] = x
        F = x.shape[-1]
        h = lstm_seq(x.reshape(B * T, N, F), p[f"note{l}.lstm.W"], p[f"note{l}.lstm.U"],
                     p[f"note{l}.lstm.b"], recurrent_activation)
        x = h.reshape(B, T, N, -1)
        if taps is not None:
            taps[f"note{l}.h"] = x
        x = dropout(x, masks.get(site[l][1]) if l < 2 else None, dropout_rate)
    pr = torch.sigmoid(x @ p["note_dense.W"] + p["note_dense.b"])     # model.py:94
    vol = x @ p["volume_dense.W"] + p["volume_dense.b"]               # model.py:95
    return torch.cat([pr, vol], dim=-1)


def model_forward(p, cfg: Config, notes_in, chosen_in, beat_in, style_in, masks=None,
                  input_dropout=0.2, dropout_rate=0.5, recurrent_activation="hard_sigmoid",
                  taps: Optional[dict] = None):
    """model.py:128-152: `model([notes, chosen, beat, style])` -> [B,T,N,3]."""
    masks = masks or {}
    notes = dropout(notes_in, masks.get("D1"), input_dropout)
    beat = dropout(beat_in, masks.get("D2"), input_dropout)
    chosen = dropout(chosen_in, masks.get("D3"), input_dropout)
    style = style_embed(p, style_in)
    if taps is not None:
        taps["style"] = style
    time_out = time_axis_forward(p, cfg, notes, beat, style, masks, dropout_rate,
                                 recurrent_activation, taps)
    if taps is not None:
        taps["time_out"] = time_out
    return note_axis_forward(p, cfg, time_out, chosen, style, masks, dropout_rate,
                             recurrent_activation, taps)


def time_model_predict(p, cfg: Config, notes, beat, style_in, batch_size=32,
                       recurrent_activation="hard_sigmoid"):
    """`time_model.predict([notes, beat, style])` (model.py:155).  Keras predict
    runs in chunks of batch_size=32; the chunk scopes the pitch_bins scramble."""
    outs = []
    for s in range(0, notes.shape[0], batch_size):
        sl = slice(s, s + batch_size)
        outs.append(time_axis_forward(p, cfg, notes[sl], beat[sl], style_embed(p, style_in[sl]),
                                      None, 0.0, recurrent_activation))
    return torch.cat(outs, dim=0)


def note_model_predict(p, cfg: Config, note_features, chosen, style_in,
                       recurrent_activation="hard_sigmoid"):
    """`note_model.predict([note_features, chosen, style])` (model.py:157-167),
    time dimension 1."""
    return note_axis_forward(p, cfg, note_features, chosen, style_embed(p, style_in),
                             None, 0.0, recurrent_activation)


# --------------------------------------------------------------------------
# training step: loss, autograd gradients, Keras-2 Nadam (model.py:152)
# --------------------------------------------------------------------------
def _f32(x: float) -> float:
    return float(np.float32(x))


@dataclass
class NadamState:
    # keras.optimizers.Nadam.__init__ stores lr / beta_1 / beta_2 as K.variable(...) of floatx = float32, so the graph
    # computes with the float32-ROUNDED values: 1 - beta_2 is 0.00099998713, not 0.001 (1.3e-5 relative, visible in
    # the first updates).  The oracle keeps its arithmetic in the caller's dtype but uses those rounded constants.
    lr: float = _f32(0.002)
    beta_1: float = _f32(0.9)
    beta_2: float = _f32(0.999)
    eps: float = 1e-8            # Keras <= 2.1.2; K.epsilon()=1e-7 from 2.1.3
    schedule_decay: float = 0.004
    iterations: int = 0
    m_schedule: float = 1.0
    m: Dict[str, torch.Tensor] = field(default_factory=dict)
    v: Dict[str, torch.Tensor] = field(default_factory=dict)


def nadam_scalars(st: NadamState):
    """The per-step scalars of keras.optimizers.Nadam.get_updates (host side)."""
    t = st.iterations + 1
    mu_t = st.beta_1 * (1.0 - 0.5 * (0.96 ** (t * st.schedule_decay)))
    mu_t1 = st.beta_1 * (1.0 - 0.5 * (0.96 ** ((t + 1) * st.schedule_decay)))
    m_sched_new = st.m_schedule * mu_t
    m_sched_next = st.m_schedule * mu_t * mu_t1
    return t, mu_t, mu_t1, m_sched_new, m_sched_next


def nadam_step(p: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor], st: NadamState):
    t, mu_t, mu_t1, ms_new, ms_next = nadam_scalars(st)
    for k in p:
        g = grads[k]
        if k not in st.m:
            st.m[k] = torch.zeros_like(p[k]); st.v[k] = torch.zeros_like(p[k])
        g_prime = g / (1.0 - ms_new)
        st.m[k] = st.beta_1 * st.m[k] + (1.0 - st.beta_1) * g
        m_prime = st.m[k] / (1.0 - ms_next)
        st.v[k] = st.beta_2 * st.v[k] + (1.0 - st.beta_2) * g * g
        v_prime = st.v[k] / (1.0 - st.beta_2 ** t)
        m_bar = (1.0 - mu_t) * g_prime + mu_t1 * m_prime
        p[k] = p[k] - st.lr * m_bar / (torch.sqrt(v_prime) + st.eps)
    st.iterations += 1
    st.m_schedule = ms_new
    return p


def loss_and_grads(p, cfg: Config, notes_in, chosen_in, beat_in, style_in, target, masks=None,
                   recurrent_activation="hard_sigmoid", input_dropout=0.2, dropout_rate=0.5):
    """One forward + primary_loss + backward; gradients by autograd (independent
    of the hand-derived CUDA backward)."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    out = model_forward(leaf, cfg, notes_in, chosen_in, beat_in, style_in, masks,
                        input_dropout, dropout_rate, recurrent_activation)
    loss = primary_loss(target, out)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaf.items()}
    return loss.detach(), out.detach(), grads


# --------------------------------------------------------------------------
# generate.py
# --------------------------------------------------------------------------
def compute_beat(t: int, notes_in_bar: int) -> np.ndarray:       # dataset.py:14-15
    v = np.zeros(notes_in_bar); v[t % notes_in_bar] = 1
    return v


def compute_genre(genre_id: int, genre_sizes: Sequence[int] = (3, 6, 14)) -> np.ndarray:
    # dataset.py:20-26 with constants.py:10-40 group sizes
    v = np.zeros(sum(genre_sizes))
    start = sum(genre_sizes[:genre_id])
    v[start:start + genre_sizes[genre_id]] = 1 / genre_sizes[genre_id]
    return v


def apply_temperature(prob: np.ndarray, temperature: float) -> np.ndarray:
    """generate.py:81-91 -- runs in the dtype of `prob` (float32 from predict)."""
    if temperature != 1:
        with np.errstate(divide="ignore", over="ignore"):
            x = -np.log(1 / prob - 1)
            prob = 1 / (1 + np.exp(-x / temperature))
    return prob


class UniformStream:
    """Stands in for the global np.random.random() stream of generate.py:52,57."""

    def __init__(self, values: np.ndarray):
        self.values = np.asarray(values, dtype=np.float64)
        self.pos = 0

    def __call__(self) -> float:
        v = self.values[self.pos]
        self.pos += 1
        return float(v)


class Generation:
    """generate.py:13-79 `MusicGeneration` (window memories as arrays)."""

    def __init__(self, cfg: Config, style: np.ndarray, default_temp: float = 1):
        self.cfg = cfg
        L, N = cfg.seq_len, cfg.num_notes
        self.notes_memory = np.zeros((L, N, cfg.note_units))          # generate.py:18
        self.beat_memory = np.zeros((L, cfg.notes_per_bar))           # :19 (zeros, not one-hot)
        self.style_memory = np.tile(np.asarray(style, dtype=np.float64), (L, 1))   # :20
        self.next_note = np.zeros((N, cfg.note_units))
        self.silent_time = cfg.notes_per_bar                          # :24
        self.default_temp = default_temp
        self.temperature = default_temp
        self.results: List[np.ndarray] = []
        self.min_margin = np.inf

    def choose(self, prob: np.ndarray, n: int, rand) -> None:         # generate.py:47-58
        vol = prob[n, -1]
        pr = apply_temperature(prob[n, :-1], self.temperature)
        u = rand()
        self.min_margin = min(self.min_margin, abs(u - float(pr[0])))
        if u <= pr[0]:
            self.next_note[n, 0] = 1
            self.next_note[n, 2] = vol
            u2 = rand()
            self.min_margin = min(self.min_margin, abs(u2 - float(pr[1])))
            if u2 <= pr[1]:
                self.next_note[n, 1] = 1

    def end_time(self, t: int) -> np.ndarray:                         # generate.py:60-79
        if np.count_nonzero(self.next_note) == 0:
            self.silent_time += 1
            if self.silent_time >= self.cfg.notes_per_bar:
                self.temperature += 0.1
        else:
            self.silent_time = 0
            self.temperature = self.default_temp
        self.notes_memory = np.concatenate([self.notes_memory[1:], self.next_note[None]], 0)
        self.beat_memory = np.concatenate(
            [self.beat_memory[1:], compute_beat(t, self.cfg.notes_per_bar)[None]], 0)
        self.results.append(self.next_note)
        self.next_note = np.zeros_like(self.next_note)
        return self.results[-1]


def generate(p, cfg: Config, styles: Sequence[np.ndarray], num_steps: int, uniforms: np.ndarray,
             mode: str = "incremental", dtype=torch.float32, recurrent_activation="hard_sigmoid",
             default_temp: float = 1, forced_events: Optional[np.ndarray] = None,
             return_probs: bool = False):
    """generate.py:98-121.  `literal` repeats the reference's call structure
    (one full-window time_model.predict + 48 full note_model.predict per
    timestep); `incremental` carries the note-axis LSTM state from note to note
    (mathematically identical because the note LSTM is causal in n and
    next_note[0..n-1] is final).  The uniform stream is consumed in reference
    order: n-major, sequence-minor, second draw only when played.

    forced_events [steps, G, N, 3]: lock-step mode -- decisions are still made
    (and margins recorded) but the state that is fed back is the forced event.
    """
    rand = UniformStream(uniforms)
    gens = [Generation(cfg, s, default_temp) for s in styles]
    G, N, Ut = len(gens), cfg.num_notes, cfg.time_axis_units
    act = recurrent_activation
    out_steps, prob_steps, decision_steps, temp_steps = [], [], [], []
    with torch.no_grad():
        for t in range(num_steps):
            temp_steps.append([g.temperature for g in gens])      # temperature in effect while step t is sampled
            notes = torch.tensor(np.stack([g.notes_memory for g in gens]), dtype=dtype)
            beat = torch.tensor(np.stack([g.beat_memory for g in gens]), dtype=dtype)
            sty = torch.tensor(np.stack([g.style_memory for g in gens]), dtype=dtype)
            feats = time_model_predict(p, cfg, notes, beat, sty, 32, act)[:, -1:, :, :]   # :108-109
            sty1 = sty[:, -1:, :]
            probs_t = np.zeros((G, N, cfg.note_units), dtype=np.float32 if dtype == torch.float32 else np.float64)
            decisions_t = np.zeros((G, N, cfg.note_units))
            if mode == "literal":
                for n in range(N):
                    chosen = torch.tensor(np.stack([g.next_note for g in gens])[:, None], dtype=dtype)
                    pred = note_model_predict(p, cfg, feats, chosen, sty1, act).numpy()
                    for i, g in enumerate(gens):
                        probs_t[i, n] = pred[i, 0, n]
                        g.choose(pred[i][-1], n, rand)
                        decisions_t[i, n] = g.next_note[n]
                        if forced_events is not None:
                            g.next_note[n] = forced_events[t, i, n]
            else:
                emb = style_embed(p, sty1)[:, 0]                                  # [G,64]
                sps = [torch.tanh(emb @ p[f"note{l}.sd.W"] + p[f"note{l}.sd.b"])
                       for l in range(cfg.note_axis_layers)]
                un = cfg.note_axis_units
                hs = [feats.new_zeros(G, un) for _ in range(cfg.note_axis_layers)]
                cs = [feats.new_zeros(G, un) for _ in range(cfg.note_axis_layers)]
                fn = hard_sigmoid if act == "hard_sigmoid" else torch.sigmoid
                for n in range(N):
                    prev = np.stack([g.next_note[n - 1] if n > 0 else np.zeros(cfg.note_units)
                                     for g in gens])
                    x = torch.cat([feats[:, 0, n, :], torch.tensor(prev, dtype=dtype)], dim=1)
                    for l in range(cfg.note_axis_layers):
                        x = x + sps[l]
                        z = x @ p[f"note{l}.lstm.W"] + p[f"note{l}.lstm.b"] + hs[l] @ p[f"note{l}.lstm.U"]
                        i_, f_ = fn(z[:, :un]), fn(z[:, un:2 * un])
                        g_, o_ = torch.tanh(z[:, 2 * un:3 * un]), fn(z[:, 3 * un:])
                        cs[l] = f_ * cs[l] + i_ * g_
                        hs[l] = o_ * torch.tanh(cs[l])
                        x = hs[l]
                    pr = torch.sigmoid(x @ p["note_dense.W"] + p["note_dense.b"])
                    vol = x @ p["volume_dense.W"] + p["volume_dense.b"]
                    pred_n = torch.cat([pr, vol], dim=1).numpy()
                    for i, g in enumerate(gens):
                        probs_t[i, n] = pred_n[i]
                        row = np.zeros((N, cfg.note_units), dtype=pred_n.dtype)
                        row[n] = pred_n[i]
                        g.choose(row, n, rand)
                        decisions_t[i, n] = g.next_note[n]
                        if forced_events is not None:
                            g.next_note[n] = forced_events[t, i, n]
            out_steps.append(np.stack([g.end_time(t) for g in gens]))
            prob_steps.append(probs_t)
            decision_steps.append(decisions_t)
    events = np.stack(out_steps)            # [steps, G, N, 3]
    info = {"uniforms_used": rand.pos, "min_margin": min(g.min_margin for g in gens)}
    info["decisions"] = np.stack(decision_steps)   # the oracle's own draws (== events unless forced)
    info["temperature_trace"] = np.array(temp_steps, dtype=np.float64)          # [steps, G]
    info["temperature"] = np.array([g.temperature for g in gens], dtype=np.float64)
    info["silent_time"] = np.array([g.silent_time for g in gens])
    if return_probs:
        info["probs"] = np.stack(prob_steps)
    return events, info


# --------------------------------------------------------------------------
# synthetic data of the constants.py shape
# --------------------------------------------------------------------------
def synthetic_batch(cfg: Config, B: int, T: Optional[int] = None, seed: int = 1234,
                    dtype=torch.float32):
    T = T or cfg.seq_len
    rs = np.random.RandomState(seed)
    N = cfg.num_notes
    roll = np.zeros((B, T + 1, N, 3))
    play = (rs.random_sample((B, T + 1, N)) < 0.05).astype(np.float64)
    roll[..., 0] = play
    roll[..., 1] = (rs.random_sample((B, T + 1, N)) < 0.1) * play
    roll[..., 2] = rs.uniform(0.2, 0.8, (B, T + 1, N)) * play
    notes, target = roll[:, :T], roll[:, 1:]
    phase = rs.randint(0, cfg.notes_per_bar, B)
    beat = np.zeros((B, T, cfg.notes_per_bar))
    tt = (np.arange(T)[None, :] + phase[:, None]) % cfg.notes_per_bar
    beat[np.arange(B)[:, None], np.arange(T)[None, :], tt] = 1
    style = np.zeros((B, T, cfg.num_styles))
    style[np.arange(B), :, rs.randint(0, cfg.num_styles, B)] = 1
    to = lambda a: torch.tensor(a, dtype=dtype)
    # model.py:151 input order is [notes, chosen(=target), beat, style]
    return to(notes), to(target), to(beat), to(style), to(target)


def random_masks(cfg: Config, B: int, T: int, seed: int = 7, input_dropout=0.2, dropout_rate=0.5):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, shp in dropout_site_shapes(cfg, B, T).items():
        if shp is None:
            continue
        rate = input_dropout if k in ("D1", "D2", "D3") else dropout_rate
        out[k] = (torch.rand(shp, generator=g) >= rate).to(torch.float32)
    return out
