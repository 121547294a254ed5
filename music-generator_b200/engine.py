"""Host-side sequencing of the DeepJ hot path on one B200.

PyTorch is used only for device memory, streams and (in trainer.py)
torch.distributed; every arithmetic step is a kernel of libdeepj_sm100.so
called through the C ABI (include/deepj_b200.h).  The engine mirrors the
reference graph of model.py:128-152 (forward), tf.gradients of it (backward)
and keras Nadam (model.py:152).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib
from ._lib import DJ_BF16, DJ_F16, DJ_F32, NO_DROPOUT, Dropout, check
from .config import ModelConfig, param_shapes, round_up

N = 48


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Workspace:
    """Activation / gradient buffers for one (B, T, precision, train) shape."""

    def __init__(self, cfg: ModelConfig, B: int, T: int, prec: str, train: bool, dev):
        self.B, self.T, self.M = B, T, B * T * N
        bf16, mixed = prec in ("bf16", "mixed"), prec == "mixed"
        gen = prec == "gen"     # generation window: fp32-grade tensor-core path (half hi+lo operands), time axis only
        M, BT = self.M, B * T
        f32 = dict(dtype=torch.float32, device=dev)
        adt = torch.bfloat16 if bf16 else torch.float16 if gen else torch.float32
        self.gen = gen
        self.h_hi, self.h_lo = [], []     # gen: h_{t-1} as half hi / lo, the exchange buffers of the inference scan
        self.emb = torch.empty(BT, cfg.style_units, **f32)
        self.sp, self.A, self.Z, self.h, self.c = [], [], [], [], []
        self.A_lo = []      # mixed: bf16 residual of A (second operand of the split gate GEMM)
        self.ld = []
        for L in cfg.layers():
            ld = round_up(L["F"], 32)
            self.ld.append(ld)
            self.sp.append(torch.empty(BT, L["F"], **f32))
            self.A.append(torch.empty(M, ld, dtype=adt, device=dev))
            self.A_lo.append(torch.empty(M, ld, dtype=adt, device=dev) if (mixed or gen) else None)
            tl = gen and L["axis"] == "time"
            # + one spare timestep of rows: the fused two-layer scan keeps h of the LAST step at row T for layer 1
            self.h_hi.append(torch.empty(M + N, L["U"], dtype=torch.float16, device=dev) if tl else None)
            self.h_lo.append(torch.empty(M + N, L["U"], dtype=torch.float16, device=dev) if tl else None)
            self.Z.append(torch.empty(M, 4 * L["U"], **f32))
            self.h.append(torch.empty(M, L["U"], **f32))
            self.c.append(torch.empty(M, L["U"], **f32) if train else None)
        # h_{step-1} in 16 bits (half in mixed precision): B operand of the tensor-core recurrence and A operand of
        # the recurrent-weight gradient
        hdt = torch.float16 if mixed else torch.bfloat16
        self.hprev = [torch.empty(M, L["U"], dtype=hdt, device=dev) if (train and bf16) else None
                      for L in cfg.layers()]
        # activated gates saved by the tensor-core forward scan for the reverse scan: 4 IEEE halves per cell
        self.G16 = [torch.empty(M, 4 * L["U"], dtype=torch.float16, device=dev) if (train and bf16) else None
                    for L in cfg.layers()]
        self.probs = torch.empty(M, 3, **f32)
        if gen:
            # fused two-layer generation scan (dj_lstm_scan_tc_gen2): per-sequence constant part of layer 1's
            # pre-activations (style projection through W1, + bias) and the per-tile step counters
            self.c1 = torch.empty(B, 4 * cfg.time_axis_units, **f32)
            self.c1_version = -1
            self.flags = torch.zeros(2 * (B * N // 16), dtype=torch.int32, device=dev)
        if train:
            un = cfg.note_axis_units
            self.dXtop = torch.empty(M, un, **f32)
            self.partials = torch.empty(_lib.load().dj_head_partials_size(un), **f32)
            self.loss = torch.zeros(1, **f32)
            # one dZ per layer: the weight-gradient GEMMs of layer l run on a second stream while the
            # reverse scan of layer l-1 is already writing its own dZ
            self.dZ = [torch.empty(M, 4 * L["U"], dtype=adt, device=dev) for L in cfg.layers()]   # bf16 (range)
            self.dA = [torch.empty(M, ld, **f32) for ld in self.ld]
            self.ds = [torch.empty(BT, L["F"], **f32) for L in cfg.layers()]
            self.demb = torch.empty(BT, cfg.style_units, **f32)


_M32 = 0xFFFFFFFF


def _mix32(x: int) -> int:
    x ^= x >> 16; x = (x * 0x7feb352d) & _M32
    x ^= x >> 15; x = (x * 0x846ca68b) & _M32
    return x ^ (x >> 16)


def site_key(seed: int, site: int) -> int:
    """dj_site_key of csrc/dj_common.cuh (the key dj_make_dropout derives from (seed, site)), on the host."""
    seed &= 0xFFFFFFFFFFFFFFFF
    k = _mix32((seed & _M32) ^ ((0x9E3779B9 * (site + 1)) & _M32))
    return _mix32((k + (seed >> 32)) & _M32)


class StepParams:
    """What changes from one training step to the next and is passed to kernels by VALUE in the eager path -- the 12
    dropout site keys of the step's seed, the ten Nadam scalars, the exchange epoch -- kept in 128 bytes of device
    memory instead, so that a captured CUDA graph of the step can be replayed: words 0..14 site keys, word 15 epoch,
    words 16..25 floats.  Uploads go through a ring of pinned host rows (the host runs ahead of the GPU)."""
    SLOTS = 64

    def __init__(self, dev):
        self.buf = torch.zeros(32, dtype=torch.int32, device=dev)
        self.pinned = torch.zeros(self.SLOTS, 32, dtype=torch.int32).pin_memory()
        self.rows = self.pinned.numpy()
        self.events = [None] * self.SLOTS
        self.i = 0
        base = self.buf.data_ptr()
        self.key_ptr = lambda site: base + 4 * site
        self.epoch_ptr, self.scalar_ptr = base + 4 * 15, base + 4 * 16

    def upload(self, seed: int, epoch: int, scalars) -> None:
        s = self.i % self.SLOTS
        self.i += 1
        if self.events[s] is not None:
            self.events[s].synchronize()          # the copy that last used this pinned row has been performed
        row = self.rows[s]
        row.view(np.uint32)[:15] = [site_key(seed, k) for k in range(15)]
        row.view(np.uint32)[15] = epoch
        row.view(np.float32)[16:26] = scalars
        self.buf.copy_(self.pinned[s], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[s] = ev


class Engine:
    def __init__(self, cfg: ModelConfig = ModelConfig(), device: Optional[torch.device] = None,
                 precision: str = "mixed", recurrent_activation: str = "hard_sigmoid",
                 input_dropout: float = 0.2, dropout: float = 0.5, deterministic: Optional[bool] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("the DeepJ B200 engine needs a CUDA device; there is no CPU fallback")
        # "mixed" (training default): tensor cores at fp32 grade where the outputs need it -- split bf16 hi+lo gate
        #     GEMM (3 passes), half-precision h and hi+lo half-precision U in the recurrence; bf16 in the backward pass.
        # "bf16": every tensor-core operand one bf16 (fastest; outside north_star's 1e-3 tolerance).
        # "fp32": CUDA-core kernels throughout (generation, reference-grade checks).
        assert precision in ("mixed", "bf16", "fp32")
        self.lib = _lib.load()
        self.cfg = cfg
        self.dev = device or torch.device("cuda", torch.cuda.current_device())
        self.precision = precision
        self.hard = 1 if recurrent_activation == "hard_sigmoid" else 0
        self.input_dropout, self.dropout = input_dropout, dropout
        self.layers = cfg.layers()
        # ---- flat parameter / gradient / optimizer buffers (one NCCL message)
        self.shapes = param_shapes(cfg)
        self.offsets: Dict[str, int] = {}
        off = 0
        for k, shp in self.shapes.items():
            self.offsets[k] = off
            off += round_up(int(np.prod(shp)), 32)
        self.flat_size = off
        self.num_params = sum(int(np.prod(s)) for s in self.shapes.values())
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.flat = torch.zeros(off, **f32)
        self.gflat = torch.zeros(off, **f32)
        self.m = torch.zeros(off, **f32)
        self.v = torch.zeros(off, **f32)
        self.params = {k: self._view(self.flat, k) for k in self.shapes}
        self.grads = {k: self._view(self.gflat, k) for k in self.shapes}
        self.iterations, self.m_schedule = 0, 1.0
        self.nadam = dict(lr=0.002, beta_1=0.9, beta_2=0.999, eps=1e-8, schedule_decay=0.004)
        self._ws: Dict[tuple, Workspace] = {}
        self._wbf: Dict[str, torch.Tensor] = {}
        self._wbf_version = -1
        self._wbf_mixed = None
        self._version = 0
        self.launches = 0   # kernels of ours enqueued (bench.py reports it)
        self.profile = None
        self.profile_only = None
        self.tc_scan = True   # tensor-core training: recurrence on tcgen05 (False: debugging, precision "bf16" only)
        # backward runs on two streams: the dependency chain (reverse scan -> data gradient -> next layer's
        # reverse scan) on a high-priority stream, the weight / style / conv gradients behind it on the caller's
        self.overlap = os.environ.get("DJ_NO_OVERLAP", "") == ""
        self._hi = None
        self._light = None
        self.bwd_streams = int(os.environ.get("DJ_BWD_STREAMS", "3"))
        self._tag = ""
        self.peer = None      # parallel.PeerNadam: fused gradient exchange + Nadam over peer memory
        # The training step as ONE CUDA graph launch (DJ_GRAPH=0: ~56 launches from Python).  Per-step values live in
        # device memory (StepParams); the first step of a shape runs eagerly through the same kernels, the second is
        # captured, every later one is a replay.
        self.graph = os.environ.get("DJ_GRAPH", "1") != "0"
        self._sp: Optional[StepParams] = None
        self._graphs: Dict[tuple, dict] = {}
        self._dev_drops = None
        # generation window on the tensor cores at fp32 grade (half hi+lo operands, 3 MMA passes); DJ_GEN_TC=0 = the
        # CUDA-core fp32 kernels
        self.gen_tc = os.environ.get("DJ_GEN_TC", "1") != "0"
        # one or two sequences: both time-axis layers of the window in ONE launch, as a wavefront (DJ_GEN_FUSED=0: one
        # inference scan per layer)
        self.gen_fused = os.environ.get("DJ_GEN_FUSED", "1") != "0"
        # Deterministic gradients (DJ_DETERMINISTIC=1): every sum over CTAs in the backward pass (split weight-gradient
        # GEMMs, bias / conv / style gradients) goes through per-CTA partials added in index order instead of fp32
        # atomics, so two runs -- and 1 GPU vs N GPUs on the same shards -- give the same bits.  One workspace per
        # stream the backward pass launches on (dj_set_reduce_workspace).
        self.deterministic = (os.environ.get("DJ_DETERMINISTIC", "0") == "1") if deterministic is None else deterministic
        self._red_ws: Dict[int, torch.Tensor] = {}

    # ------------------------------------------------------------------ params
    def _view(self, flat, k):
        shp = self.shapes[k]
        n = int(np.prod(shp))
        return flat[self.offsets[k]:self.offsets[k] + n].view(*shp)

    def rebind_flat(self, flat: torch.Tensor, gflat: torch.Tensor) -> None:
        """Move the flat parameter / gradient buffers into caller-provided device memory (parallel.PeerNadam puts
        them into an allocation the other ranks can map); the current weights are carried over."""
        assert flat.numel() == self.flat_size and gflat.numel() == self.flat_size
        flat.copy_(self.flat)
        gflat.zero_()
        self.flat, self.gflat = flat, gflat
        self.params = {k: self._view(self.flat, k) for k in self.shapes}
        self.grads = {k: self._view(self.gflat, k) for k in self.shapes}
        self._version += 1

    # The LSTM tensors live on the device in GATE-INTERLEAVED column order
    # (col = 4*unit + gate) so a cell's four gates are one 16-byte vector; the
    # Keras block order [i | f | c | o] exists only at this API boundary.
    @staticmethod
    def _is_lstm(k: str) -> bool:
        return ".lstm." in k

    @staticmethod
    def _to_internal(v: torch.Tensor) -> torch.Tensor:
        u = v.shape[-1] // 4
        return v.reshape(*v.shape[:-1], 4, u).transpose(-1, -2).reshape(v.shape).contiguous()

    @staticmethod
    def _to_keras(v: torch.Tensor) -> torch.Tensor:
        u = v.shape[-1] // 4
        return v.reshape(*v.shape[:-1], u, 4).transpose(-1, -2).reshape(v.shape).contiguous()

    def set_params(self, state: Dict[str, "np.ndarray | torch.Tensor"]) -> None:
        for k in self.shapes:
            v = torch.as_tensor(np.asarray(state[k]) if not torch.is_tensor(state[k]) else state[k])
            if tuple(v.shape) != tuple(self.shapes[k]):
                raise ValueError(f"{k}: expected shape {self.shapes[k]}, got {tuple(v.shape)}")
            v = v.to(torch.float32)
            self.params[k].copy_(self._to_internal(v) if self._is_lstm(k) else v)
        self._version += 1

    def _export(self, views) -> Dict[str, np.ndarray]:
        out = {}
        for k, v in views.items():
            v = v.detach().cpu()
            out[k] = (self._to_keras(v) if self._is_lstm(k) else v).numpy().copy()
        return out

    def get_params(self) -> Dict[str, np.ndarray]:
        """The 28 tensors in Keras layouts."""
        return self._export(self.params)

    def get_grads(self) -> Dict[str, np.ndarray]:
        """Gradients of the last backward() in Keras layouts."""
        return self._export(self.grads)

    def init_params(self, seed: int = 0) -> None:
        """Keras default initialisers (glorot_uniform / orthogonal / zeros, forget bias 1)."""
        g = torch.Generator().manual_seed(seed)
        st = {}
        for name, shp in self.shapes.items():
            if name.endswith(".b"):
                t = torch.zeros(shp, dtype=torch.float64)
                if ".lstm." in name:
                    u = shp[0] // 4
                    t[u:2 * u] = 1.0
            elif name.endswith("lstm.U"):
                u = shp[0]
                blocks = []
                for _ in range(4):
                    q, r = torch.linalg.qr(torch.randn(u, u, generator=g, dtype=torch.float64))
                    blocks.append(q * torch.sign(torch.diagonal(r)))
                t = torch.cat(blocks, dim=1) * 0.5
            else:
                fan_in, fan_out = (shp[0] * shp[1], shp[0] * shp[2]) if len(shp) == 3 else shp
                lim = math.sqrt(6.0 / (fan_in + fan_out))
                t = (torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * lim
            st[name] = t.to(torch.float32)
        self.set_params(st)

    def _call(self, name, *args):
        if self.profile is not None and (self.profile_only is None or (name + self._tag) in self.profile_only):
            # per-entry-point device time: CUDA events on the launching stream around the launch
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            check(getattr(self.lib, name)(*args), name)
            e1.record()
            self.profile.append((name + self._tag, e0, e1))
        else:
            check(getattr(self.lib, name)(*args), name)
        self.launches += 1

    def profile_summary(self):
        torch.cuda.synchronize()
        agg = {}
        for name, e0, e1 in self.profile:
            t = e0.elapsed_time(e1)
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1; a[1] += t
        self.profile = []
        return agg

    def _refresh_bf16(self, mixed: bool):
        """16-bit operand copies of the LSTM kernels: transposed [4U, ld] (forward B operand, K-major) and natural
        [F, 4U] (data-gradient B operand).  Mixed precision adds the residuals W^T - bf16(W^T) and keeps the recurrent
        kernel of the forward scan as IEEE half hi + lo."""
        if self._wbf_version == self._version and self._wbf_mixed == mixed:
            return
        ents = []   # (src, rows, cols, dst, dst_lo, ldo, transpose, fmt)
        rfmt = DJ_F16 if mixed else DJ_BF16
        for li, L in enumerate(self.layers):
            W = self.params[f"{L['name']}.lstm.W"]
            Um = self.params[f"{L['name']}.lstm.U"]
            F, U4 = W.shape
            U, ld = L["U"], round_up(F, 32)
            n = L["name"]
            if f"{n}.Wt" not in self._wbf:
                w16 = dict(dtype=torch.int16, device=self.dev)     # raw 16-bit words: bf16 or half by mode
                self._wbf[f"{n}.Wt"] = torch.empty(U4, ld, **w16)     # W^T  [4U, ld]: forward B operand (K-major)
                self._wbf[f"{n}.Wt_lo"] = torch.empty(U4, ld, **w16)  #       its bf16 residual (split gate GEMM)
                self._wbf[f"{n}.Wn"] = torch.empty(F, U4, **w16)      # W    [F, 4U]: data-gradient B operand
                self._wbf[f"{n}.Ut"] = torch.empty(U4, U, **w16)      # U^T  [4U, U]: resident A operand, forward scan
                self._wbf[f"{n}.Ut_lo"] = torch.empty(U4, U, **w16)   #       its residual (second MMA pass)
                self._wbf[f"{n}.Un"] = torch.empty(U, U4, **w16)      # U    [U, 4U]: resident A operand, reverse scan
            lo = (lambda k: self._wbf[k]) if mixed else (lambda k: None)
            ents += [(W, F, U4, self._wbf[f"{n}.Wt"], lo(f"{n}.Wt_lo"), ld, 1, DJ_BF16),
                     (W, F, U4, self._wbf[f"{n}.Wn"], None, U4, 0, DJ_BF16),
                     (Um, U, U4, self._wbf[f"{n}.Ut"], lo(f"{n}.Ut_lo"), U, 1, rfmt),
                     (Um, U, U4, self._wbf[f"{n}.Un"], None, U4, 0, DJ_BF16)]
        k = len(ents)
        ints = lambda j: (C.c_int * k)(*[e[j] for e in ents])
        self._call("dj_cast16_multi", k, (C.c_void_p * k)(*[e[0].data_ptr() for e in ents]), ints(1), ints(2),
                   (C.c_void_p * k)(*[e[3].data_ptr() for e in ents]),
                   (C.c_void_p * k)(*[None if e[4] is None else e[4].data_ptr() for e in ents]),
                   ints(5), ints(6), ints(7), None, _stream())
        self._wbf_version, self._wbf_mixed = self._version, mixed

    GEN_SCALE = 1024.0      # power of two: keeps the half-precision residuals of W and U out of the subnormal range

    def _refresh_gen(self):
        """Generation window on the tensor cores at fp32 grade: the time-axis kernels W^T and U^T as IEEE half hi + lo
        (~22 mantissa bits), pre-scaled by GEN_SCALE (undone exactly in the kernels' epilogues)."""
        if getattr(self, "_wgen_version", -1) == self._version:
            return
        ents = []
        for L in self.layers:
            if L["axis"] != "time":
                continue
            n = L["name"]
            W, Um = self.params[f"{n}.lstm.W"], self.params[f"{n}.lstm.U"]
            F, U4 = W.shape
            U, ld = L["U"], round_up(F, 32)
            if f"{n}.gWt" not in self._wbf:
                w16 = dict(dtype=torch.float16, device=self.dev)
                for k, shp in (("gWt", (U4, ld)), ("gWt_lo", (U4, ld)), ("gUt", (U4, U)), ("gUt_lo", (U4, U))):
                    self._wbf[f"{n}.{k}"] = torch.empty(*shp, **w16)
            ents += [(W, F, U4, self._wbf[f"{n}.gWt"], self._wbf[f"{n}.gWt_lo"], ld, 1, DJ_F16),
                     (Um, U, U4, self._wbf[f"{n}.gUt"], self._wbf[f"{n}.gUt_lo"], U, 1, DJ_F16)]
        k = len(ents)
        ints = lambda j: (C.c_int * k)(*[e[j] for e in ents])
        self._call("dj_cast16_multi", k, (C.c_void_p * k)(*[e[0].data_ptr() for e in ents]), ints(1), ints(2),
                   (C.c_void_p * k)(*[e[3].data_ptr() for e in ents]), (C.c_void_p * k)(*[e[4].data_ptr() for e in ents]),
                   ints(5), ints(6), ints(7), (C.c_float * k)(*[self.GEN_SCALE] * k), _stream())
        self._wgen_version = self._version

    # --------------------------------------------------------------- dropout
    def _drops(self, train: bool, seed) -> Dict[int, Dropout]:
        if not train:
            return {s: NO_DROPOUT for s in range(1, 13)}
        if seed is None:        # graph path: the site keys are read from device memory (StepParams)
            if self._dev_drops is None:
                self._dev_drops = {}
                for s in range(1, 13):
                    rate = self.input_dropout if s <= 3 else self.dropout
                    d = _lib.make_dropout(0, s, rate) if rate > 0 else NO_DROPOUT
                    if rate > 0:
                        d.key_ptr = self._sp.key_ptr(s)
                    self._dev_drops[s] = d
            return self._dev_drops
        out = {}
        for s in range(1, 13):
            rate = self.input_dropout if s <= 3 else self.dropout
            out[s] = _lib.make_dropout(seed, s, rate) if rate > 0 else NO_DROPOUT
        return out

    def materialize_masks(self, B: int, T: int, seed: int) -> Dict[str, torch.Tensor]:
        """Debug: the keep masks the kernels will use for (seed), in the oracle's
        D1..D12 shapes, so a CPU checker can replay the same dropout."""
        cfg, d = self.cfg, self._drops(True, seed)
        M = B * T * N
        shapes = {1: (M, 3), 2: (B * T, cfg.notes_per_bar), 3: (M, 3), 4: (M, cfg.octave_units)}
        for L in self.layers:
            shapes[L["site_sp"]] = (M, L["F"])
            shapes[L["site_out"]] = (M, L["U"])
        out = {}
        for s, (rows, F) in shapes.items():
            t = torch.empty(rows, F, dtype=torch.float32, device=self.dev)
            self._call("dj_dropout_mask_materialize", d[s], rows, F, _ptr(t), _stream())
            out[f"D{s}"] = t.view(B, T, F) if s == 2 else t.view(B, T, N, F)
        return out

    # --------------------------------------------------------------- forward
    def workspace(self, B: int, T: int, prec: str, train: bool) -> Workspace:
        key = (B, T, prec, train)
        if key not in self._ws:
            self._ws[key] = Workspace(self.cfg, B, T, prec, train, self.dev)
        return self._ws[key]

    def _style(self, ws: Workspace, style, bstride, tstride, B, T):
        cfg, P = self.cfg, self.params
        n = len(self.layers)
        Wsd = (C.c_void_p * n)(*[P[f"{L['name']}.sd.W"].data_ptr() for L in self.layers])
        bsd = (C.c_void_p * n)(*[P[f"{L['name']}.sd.b"].data_ptr() for L in self.layers])
        Fs = (C.c_int * n)(*[L["F"] for L in self.layers])
        sp = (C.c_void_p * n)(*[t.data_ptr() for t in ws.sp])
        self._call("dj_style_fwd", _ptr(style), bstride, tstride, cfg.num_styles, B, T, _ptr(P["style.W"]),
                   _ptr(P["style.b"]), n, Wsd, bsd, Fs, _ptr(ws.emb), sp, _stream())

    def _gate_gemm(self, li: int, ws: Workspace, bf16: bool, M: int):
        L, P = self.layers[li], self.params
        U4, ld = 4 * L["U"], ws.ld[li]
        bias = P[f"{L['name']}.lstm.b"]
        self._tag = ":" + L["name"]
        if ws.gen and L["axis"] == "time":      # half hi+lo operands, weights pre-scaled by GEN_SCALE
            self._call("dj_gate_gemm_16s", _ptr(ws.A[li]), _ptr(ws.A_lo[li]), DJ_F16, ld,
                       _ptr(self._wbf[f"{L['name']}.gWt"]), _ptr(self._wbf[f"{L['name']}.gWt_lo"]), DJ_F16, ld,
                       _ptr(ws.Z[li]), U4, _ptr(bias), 1.0 / self.GEN_SCALE, M, U4, ld, _stream())
        elif bf16:
            mixed = ws.A_lo[li] is not None     # split product: A.W + A_lo.W + A.W_lo on the same TMEM accumulator
            self._call("dj_gate_gemm_16", _ptr(ws.A[li]), _ptr(ws.A_lo[li]), DJ_BF16, ld,
                       _ptr(self._wbf[f"{L['name']}.Wt"]), _ptr(self._wbf[f"{L['name']}.Wt_lo"]) if mixed else None,
                       DJ_BF16, ld, _ptr(ws.Z[li]), U4, _ptr(bias), M, U4, ld, _stream())
        else:
            self._call("dj_gemm_simt", _ptr(ws.A[li]), DJ_F32, ld, 1, _ptr(P[f"{L['name']}.lstm.W"]), DJ_F32,
                       U4, 1, _ptr(ws.Z[li]), U4, _ptr(bias), M, U4, L["F"], 0, 0, 0, _stream())

    def _tc_ok(self, B: int, T: int) -> bool:
        """The tensor-core scans work on whole tiles: 48 time-axis sequences (always whole: B*48) and 64 note-axis
        sequences (B*T % 64 == 0: any batch size at the model's 128-step windows).  There is no second code path:
        training any other shape in a tensor-core precision is an error (`tc_scan = False` is a debugging switch
        that runs the bf16 GEMMs with the CUDA-core scans)."""
        if not self.tc_scan:
            return False
        if (B * T) % 64 != 0:
            raise ValueError(f"tensor-core training needs batch*time_steps to be a multiple of 64 (got {B}*{T}); "
                             "use precision='fp32' for other shapes")
        return True

    def _scan_map(self, axis: str, B: int, T: int):
        if axis == "time":   # sequences (b,n), steps t
            return dict(S=B * N, steps=T, inner=N, outer=T * N, inner_stride=1, step=N)
        return dict(S=B * T, steps=N, inner=1, outer=N, inner_stride=0, step=1)   # sequences (b,t), steps n

    def _scan_fwd(self, li: int, ws: Workspace, B: int, T: int, train: bool):
        L = self.layers[li]
        m = self._scan_map(L["axis"], B, T)
        self._tag = ":" + L["name"]
        if ws.gen and L["axis"] == "time":
            self._call("dj_lstm_scan_tc_infer", _ptr(ws.Z[li]), _ptr(ws.h[li]), _ptr(ws.h_hi[li]), _ptr(ws.h_lo[li]),
                       _ptr(self._wbf[f"{L['name']}.gUt"]), _ptr(self._wbf[f"{L['name']}.gUt_lo"]), 1.0 / self.GEN_SCALE,
                       m["S"], m["steps"], L["U"], m["inner"], m["outer"], m["inner_stride"], m["step"], self.hard,
                       _stream())
        elif train and ws.hprev[li] is not None and self._tc_ok(B, T):
            mixed = ws.A_lo[li] is not None
            self._call("dj_lstm_scan_tc_fwd", _ptr(ws.Z[li]), _ptr(ws.G16[li]), _ptr(ws.h[li]), _ptr(ws.c[li]), _ptr(ws.hprev[li]),
                       _ptr(self._wbf[f"{L['name']}.Ut"]), _ptr(self._wbf[f"{L['name']}.Ut_lo"]) if mixed else None,
                       DJ_F16 if mixed else DJ_BF16, m["S"], m["steps"], L["U"], m["inner"], m["outer"],
                       m["inner_stride"], m["step"], self.hard, _stream())
        else:
            self._call("dj_lstm_scan_fwd", _ptr(ws.Z[li]), _ptr(ws.h[li]), _ptr(ws.c[li]) if train else None,
                       _ptr(ws.hprev[li]) if train else None,
                       _ptr(self.params[f"{L['name']}.lstm.U"]), m["S"], m["steps"], L["U"], m["inner"], m["outer"],
                       m["inner_stride"], m["step"], self.hard, _stream())
        self._tag = ""

    def forward_time(self, ws: Workspace, notes, notes_bstride, beat, beat_bstride, B, T, d, bf16, train,
                     style_done: bool = False, style=None, style_bstride=0, style_tstride=0):
        """time_axis of model.py:51-89 -> ws.h[1] ([M, Ut] canonical rows)."""
        P = self.params
        adt = DJ_BF16 if bf16 else DJ_F16 if ws.gen else DJ_F32
        if bf16:
            self._refresh_bf16(ws.A_lo[0] is not None)
        if ws.gen:
            self._refresh_gen()
        if not style_done:
            self._style(ws, style, style_bstride, style_tstride, B, T)
        self._call("dj_frontend_fwd", _ptr(notes), notes_bstride, _ptr(beat), beat_bstride, B, T,
                   _ptr(P["conv.W"]), _ptr(P["conv.b"]), _ptr(ws.sp[0]), d[1], d[2], d[4], d[5], _ptr(ws.A[0]),
                   _ptr(ws.A_lo[0]), ws.ld[0], adt, _stream())
        M = B * T * N
        self._gate_gemm(0, ws, bf16, M)
        if (ws.gen and self.gen_fused and style_done and B <= 2 and self.layers[0]["U"] == 256
                and self.layers[1]["U"] == 256):      # style_done: the generation loop, style constant over the window
            return self._scan_gen2(ws, B, T)
        self._scan_fwd(0, ws, B, T, train)
        L1 = self.layers[1]
        self._call("dj_layer_input", _ptr(ws.h[0]), self.layers[0]["U"], 0, T * N, d[6], _ptr(ws.sp[1]), L1["F"],
                   d[7], None, 0, NO_DROPOUT, B, T, _ptr(ws.A[1]), _ptr(ws.A_lo[1]), ws.ld[1], adt, _stream())
        self._gate_gemm(1, ws, bf16, M)
        self._scan_fwd(1, ws, B, T, train)

    def _scan_gen2(self, ws: Workspace, B: int, T: int):
        """Both time-axis layers of a generation window in one launch (layer 1 one step behind layer 0; its input
        projection folded into its recurrent MMA): model.py:75-85 at inference.  ws.h[1] gets the last step only."""
        L0, L1 = self.layers[0], self.layers[1]
        P, n0, n1 = self.params, L0["name"], L1["name"]
        if ws.c1_version != self._version:
            # c1[b] = sp1[b].W1 + b1: the style term of layer 1's input (constant over the window) through its kernel
            self._call("dj_gemm_simt", _ptr(ws.sp[1]), DJ_F32, T * L1["F"], 1, _ptr(P[f"{n1}.lstm.W"]), DJ_F32,
                       4 * L1["U"], 1, _ptr(ws.c1), 4 * L1["U"], _ptr(P[f"{n1}.lstm.b"]), B, 4 * L1["U"], L1["F"],
                       0, 0, 0, _stream())
            ws.c1_version = self._version
        self._tag = ":time01"
        self._call("dj_lstm_scan_tc_gen2", _ptr(ws.Z[0]), _ptr(ws.Z[1]), _ptr(ws.c1), _ptr(ws.h[1]), _ptr(ws.h_hi[0]), _ptr(ws.h_lo[0]),
                   _ptr(ws.h_hi[1]), _ptr(ws.h_lo[1]), _ptr(self._wbf[f"{n0}.gUt"]), _ptr(self._wbf[f"{n0}.gUt_lo"]),
                   _ptr(self._wbf[f"{n1}.gUt"]), _ptr(self._wbf[f"{n1}.gUt_lo"]), _ptr(self._wbf[f"{n1}.gWt"]),
                   _ptr(self._wbf[f"{n1}.gWt_lo"]), 1.0 / self.GEN_SCALE, _ptr(ws.flags), B * N, T, self.hard, _stream())
        self._tag = ""

    def forward_note(self, ws: Workspace, h_time, h_row0, h_b_rows, chosen, chosen_bstride, B, T, d, bf16,
                     train, target=None):
        """note_axis of model.py:91-126 + heads (+ loss partials when target is given)."""
        P, cfg = self.params, self.cfg
        adt = DJ_BF16 if bf16 else DJ_F32
        if bf16:
            self._refresh_bf16(ws.A_lo[2] is not None)
        M = B * T * N
        L2, L3 = self.layers[2], self.layers[3]
        self._call("dj_layer_input", _ptr(h_time), cfg.time_axis_units, h_row0, h_b_rows, d[8], _ptr(ws.sp[2]),
                   L2["F"], d[9], _ptr(chosen), chosen_bstride, d[3], B, T, _ptr(ws.A[2]), _ptr(ws.A_lo[2]), ws.ld[2], adt,
                   _stream())
        self._gate_gemm(2, ws, bf16, M)
        self._scan_fwd(2, ws, B, T, train)
        self._call("dj_layer_input", _ptr(ws.h[2]), L2["U"], 0, T * N, d[10], _ptr(ws.sp[3]), L3["F"], d[11],
                   None, 0, NO_DROPOUT, B, T, _ptr(ws.A[3]), _ptr(ws.A_lo[3]), ws.ld[3], adt, _stream())
        self._gate_gemm(3, ws, bf16, M)
        self._scan_fwd(3, ws, B, T, train)
        self._call("dj_head_loss", _ptr(ws.h[3]), cfg.note_axis_units, d[12], _ptr(P["note_dense.W"]),
                   _ptr(P["note_dense.b"]), _ptr(P["volume_dense.W"]), _ptr(P["volume_dense.b"]),
                   _ptr(target), _ptr(ws.probs), _ptr(ws.dXtop) if target is not None else None,
                   _ptr(ws.partials) if target is not None else None, M, _stream())

    def forward(self, notes, chosen, beat, style, target=None, train=False, seed=0, precision=None):
        """`model([notes, chosen, beat, style])` of model.py:151 on device tensors
        [B,T,48,3], [B,T,48,3], [B,T,16], [B,T,23] (fp32, contiguous).  Returns
        the workspace; ws.probs is the [B*T*48, 3] output."""
        prec = precision or self.precision
        bf16 = prec in ("bf16", "mixed")         # tensor-core path (16-bit operands)
        if prec == "mixed" and not self.tc_scan:
            raise ValueError("tc_scan = False (CUDA-core scans under the tensor-core GEMMs) exists for precision 'bf16' only")
        B, T = notes.shape[0], notes.shape[1]
        ws = self.workspace(B, T, prec, target is not None)
        d = self._drops(train, seed)
        self._last = dict(ws=ws, d=d, bf16=bf16, prec=prec, notes=notes, style=style, B=B, T=T)
        self.forward_time(ws, notes, T * N * 3, beat, T * 16, B, T, d, bf16, target is not None, style=style,
                          style_bstride=T * self.cfg.num_styles, style_tstride=self.cfg.num_styles)
        self.forward_note(ws, ws.h[1], 0, T * N, chosen, T * N * 3, B, T, d, bf16, target is not None, target)
        return ws

    # -------------------------------------------------------------- backward
    RED_WS_FLOATS = 8 << 20     # 32 MB: the largest user is the split weight-gradient GEMM (splits x K x 4U partial tiles)

    def _register_reduce_ws(self, stream: "torch.cuda.Stream") -> None:
        """Deterministic mode: give the library a partial-sum workspace for launches on `stream` (idempotent)."""
        h = stream.cuda_stream
        if h not in self._red_ws:
            self._red_ws[h] = torch.empty(self.RED_WS_FLOATS, dtype=torch.float32, device=self.dev)
        check(self.lib.dj_set_reduce_workspace(C.c_void_p(h), _ptr(self._red_ws[h]), self.RED_WS_FLOATS),
              "dj_set_reduce_workspace")

    def _unregister_reduce_ws(self) -> None:
        for h in self._red_ws:
            check(self.lib.dj_set_reduce_workspace(C.c_void_p(h), None, 0), "dj_set_reduce_workspace")

    def backward(self):
        """tf.gradients of primary_loss w.r.t. the 28 weight tensors, for the last
        forward(target=...).  Fills self.gflat (un-averaged local gradient) and
        ws.loss."""
        st = self._last
        ws, d, bf16, B, T = st["ws"], st["d"], st["bf16"], st["B"], st["T"]
        P, G, cfg = self.params, self.grads, self.cfg
        M, BT = B * T * N, B * T
        zdt = DJ_BF16 if bf16 else DJ_F32
        self.gflat.zero_()
        self._call("dj_head_finalize", _ptr(ws.partials), cfg.note_axis_units, _ptr(ws.loss),
                   _ptr(G["note_dense.W"]), _ptr(G["note_dense.b"]), _ptr(G["volume_dense.W"]),
                   _ptr(G["volume_dense.b"]), _stream())
        main = torch.cuda.current_stream()
        two = self.overlap and (self.profile is None or self.profile_only is not None)   # full profiling serialises
        if two and self._hi is None:
            self._hi = torch.cuda.Stream(device=self.dev, priority=-1)
        chain = self._hi if two else main
        # third stream (DJ_BWD_STREAMS=2 turns it off): the style / conv gradients (HBM- and FMA-bound, small) no longer
        # queue behind the weight-gradient GEMMs, which run in the shadow of the chain's scans and are stretched by
        # them; both side streams then have less left to do when the chain ends
        three = two and self.bwd_streams >= 3
        if three and self._light is None:
            self._light = torch.cuda.Stream(device=self.dev)
        light = self._light if three else main
        if self.deterministic:
            self._register_reduce_ws(main)
            if two:
                self._register_reduce_ws(chain)
            if three:
                self._register_reduce_ws(light)
        if two:
            chain.wait_stream(main)           # forward, zeroed gradients
        if three:
            light.wait_stream(main)
        dY, ldY = ws.dXtop, cfg.note_axis_units
        first_style = True
        if bf16 and ws.hprev[0].dtype == torch.float16:
            # the recurrence kept h_{step-1} in half; dZ is bf16 and tcgen05 kind::f16 cannot mix the two formats, so
            # the buffers are converted in place (the forward pass is done with them) -- all four now, while this
            # stream would otherwise wait for the first reverse scan, instead of one in front of each dU GEMM
            for li in range(len(self.layers)):
                self._call("dj_half_to_bf16_inplace", _ptr(ws.hprev[li]), M * self.layers[li]["U"], _stream())
        for li in (3, 2, 1, 0):
            L = self.layers[li]
            name, U, F, ld = L["name"], L["U"], L["F"], ws.ld[li]
            U4 = 4 * U
            self._tag = ":bwd:" + name
            m = self._scan_map(L["axis"], B, T)
            dZ = ws.dZ[li]
            with torch.cuda.stream(chain):
                # ---- critical chain: reverse scan, then the data gradient the next layer's scan consumes
                if bf16 and self._tc_ok(B, T):
                    self._call("dj_lstm_scan_tc_bwd", _ptr(ws.G16[li]), _ptr(ws.c[li]), _ptr(dY), ldY, d[L["site_out"]],
                               _ptr(self._wbf[f"{name}.Un"]), _ptr(dZ), _ptr(G[f"{name}.lstm.b"]), m["S"], m["steps"],
                               U, m["inner"], m["outer"], m["inner_stride"], m["step"], self.hard, _stream())
                else:
                    self._call("dj_lstm_scan_bwd", _ptr(ws.Z[li]), _ptr(ws.c[li]), _ptr(dY), ldY, d[L["site_out"]],
                               _ptr(P[f"{name}.lstm.U"]), _ptr(dZ), zdt, _ptr(G[f"{name}.lstm.b"]), m["S"], m["steps"],
                               U, m["inner"], m["outer"], m["inner_stride"], m["step"], self.hard, _stream())
                ev_scan = torch.cuda.Event() if two else None
                if two:
                    ev_scan.record(chain)
                # data gradient dA = dZ . W^T
                if bf16:
                    self._call("dj_gate_gemm_bf16", _ptr(dZ), U4, _ptr(self._wbf[f"{name}.Wn"]), U4, _ptr(ws.dA[li]),
                               ld, None, M, F, U4, _stream())
                else:
                    self._call("dj_gemm_simt", _ptr(dZ), DJ_F32, U4, 1, _ptr(P[f"{name}.lstm.W"]), DJ_F32, 1, U4,
                               _ptr(ws.dA[li]), ld, None, M, F, U4, 0, 0, 0, _stream())
                ev_dgrad = torch.cuda.Event() if two else None
                if two:
                    ev_dgrad.record(chain)
            # ---- off the chain (caller's stream): weight gradients dW = A^T.dZ, dU = H_{step-1}^T.dZ
            if two:
                main.wait_event(ev_scan)
            if bf16:
                self._call("dj_wgrad_gemm_bf16", _ptr(ws.A[li]), ld, _ptr(dZ), U4, _ptr(G[f"{name}.lstm.W"]), U4,
                           F, U4, M, _stream())
                self._call("dj_wgrad_gemm_bf16", _ptr(ws.hprev[li]), U, _ptr(dZ), U4, _ptr(G[f"{name}.lstm.U"]),
                           U4, U, U4, M, _stream())
            else:
                self._call("dj_gemm_simt", _ptr(ws.A[li]), DJ_F32, 1, ld, _ptr(dZ), zdt, U4, 1,
                           _ptr(G[f"{name}.lstm.W"]), U4, None, F, U4, M, 1, 0, 0, _stream())
                shift, period = (N, T * N) if L["axis"] == "time" else (1, N)
                self._call("dj_gemm_simt", _ptr(ws.h[li]), DJ_F32, 1, U, _ptr(dZ), zdt, U4, 1,
                           _ptr(G[f"{name}.lstm.U"]), U4, None, U, U4, M, 1, shift, period, _stream())
            # style projection backward (model.py:77-82 / 113-117) needs dA of this layer
            if two:
                light.wait_event(ev_dgrad)
            with torch.cuda.stream(light):
                self._call("dj_style_bwd_reduce", _ptr(ws.dA[li]), ld, F, _ptr(ws.sp[li]), d[L["site_sp"]], BT,
                           _ptr(ws.ds[li]), _stream())
                self._call("dj_gemm_simt", _ptr(ws.emb), DJ_F32, 1, cfg.style_units, _ptr(ws.ds[li]), DJ_F32, F, 1,
                           _ptr(G[f"{name}.sd.W"]), F, None, cfg.style_units, F, BT, 1, 0, 0, _stream())
                self._call("dj_colsum", _ptr(ws.ds[li]), F, BT, F, _ptr(G[f"{name}.sd.b"]), 1, _stream())
                self._call("dj_gemm_simt", _ptr(ws.ds[li]), DJ_F32, F, 1, _ptr(P[f"{name}.sd.W"]), DJ_F32, 1, F,
                           _ptr(ws.demb), cfg.style_units, None, BT, cfg.style_units, F, 0 if first_style else 1, 0, 0,
                           _stream())
            first_style = False
            dY, ldY = ws.dA[li], ld
        self._tag = ""
        with torch.cuda.stream(light):
            self._call("dj_conv_bwd", _ptr(st["notes"]), T * N * 3, B, T, _ptr(P["conv.W"]), _ptr(P["conv.b"]), d[1],
                       d[4], _ptr(ws.dA[0]), ws.ld[0], _ptr(G["conv.W"]), _ptr(G["conv.b"]), _stream())
            ns = cfg.num_styles
            self._call("dj_gemm_simt", _ptr(st["style"]), DJ_F32, 1, ns, _ptr(ws.demb), DJ_F32, cfg.style_units, 1,
                       _ptr(G["style.W"]), cfg.style_units, None, ns, cfg.style_units, BT, 1, 0, 0, _stream())
            self._call("dj_colsum", _ptr(ws.demb), cfg.style_units, BT, cfg.style_units, _ptr(G["style.b"]), 1, _stream())
        if two:
            main.wait_stream(chain)
        if three:
            main.wait_stream(light)
        if self.deterministic:
            self._unregister_reduce_ws()      # the registration is per calling thread and stream: leave none behind
        return ws.loss

    # ------------------------------------------------------------- optimizer
    def _nadam_scalars(self):
        """Advance the Nadam schedule by one iteration; returns the scalar arguments of the update kernels
        (lr, beta_1, beta_2, eps, mu_t, mu_t1, m_schedule_new, m_schedule_next, 1-beta_2^t)."""
        o = self.nadam
        t = self.iterations + 1
        # Keras holds lr / beta_1 / beta_2 as float32 variables: the schedule and the bias correction are functions of
        # the ROUNDED values (1 - float32(0.999)^t, not 1 - 0.999^t: 1.3e-5 apart at t = 1)
        b1, b2 = float(np.float32(o["beta_1"])), float(np.float32(o["beta_2"]))
        mu_t = b1 * (1.0 - 0.5 * (0.96 ** (t * o["schedule_decay"])))
        mu_t1 = b1 * (1.0 - 0.5 * (0.96 ** ((t + 1) * o["schedule_decay"])))
        ms_new = self.m_schedule * mu_t
        ms_next = self.m_schedule * mu_t * mu_t1
        self.iterations, self.m_schedule = t, ms_new
        self._version += 1
        return (o["lr"], b1, b2, o["eps"], mu_t, mu_t1, ms_new, ms_next, 1.0 - b2 ** t)

    def nadam_step(self, gscale: float = 1.0):
        """keras.optimizers.Nadam.get_updates on the flat buffers (model.py:152)."""
        sc = self._nadam_scalars()
        self._call("dj_nadam_step", _ptr(self.flat), _ptr(self.gflat), _ptr(self.m), _ptr(self.v), self.flat_size,
                   gscale, *sc, _stream())

    def _step_body_dev(self, gs: dict):
        """forward + backward + update with every per-step value read from device memory (what the graph captures)."""
        x = gs["inputs"]
        self.forward(x[0], x[1], x[2], x[3], target=x[4], train=True, seed=None)
        loss = self.backward()
        if self.peer is not None:
            self.peer.step_dev(self, self._sp.epoch_ptr, self._sp.scalar_ptr, _stream())
        else:
            self._call("dj_nadam_step_dev", _ptr(self.flat), _ptr(self.gflat), _ptr(self.m), _ptr(self.v), self.flat_size,
                       C.c_void_p(self._sp.scalar_ptr), _stream())
        return loss

    def _train_step_graph(self, tensors, seed: int, world: int):
        B, T = tensors[0].shape[0], tensors[0].shape[1]
        if self._sp is None:
            self._sp = StepParams(self.dev)
        key = (B, T, self.precision, self.peer is not None)
        gs = self._graphs.get(key)
        if gs is None:
            gs = self._graphs[key] = dict(inputs=[torch.empty_like(t) for t in tensors], graph=None, calls=0, launches=0)
        for dst, src in zip(gs["inputs"], tensors):
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        lr, b1, b2, eps, mu_t, mu_t1, ms_new, ms_next, bias2 = self._nadam_scalars()
        f = np.float32          # the reciprocals exactly as the C wrappers of the by-value entry points form them
        scal = [f(1.0 / world), f(lr), f(b1), f(b2), f(eps), f(mu_t), f(mu_t1), f(1) / (f(1) - f(ms_new)),
                f(1) / (f(1) - f(ms_next)), f(1) / f(bias2)]
        self._sp.upload(seed, self.peer.next_epoch() if self.peer is not None else 0, scal)
        if gs["graph"] is None and gs["calls"] >= 1:
            l0 = self.launches
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    gs["loss"] = self._step_body_dev(gs)
                gs["graph"], gs["launches"] = g, self.launches - l0
                self.launches = l0
            except Exception as e:   # the launch mechanism, not the arithmetic: log once and keep launching from Python
                print(f"[deepj] CUDA graph capture of the training step failed ({e}); launching eagerly", flush=True)
                self.graph = False
                self._dev_drops = None
        gs["calls"] += 1
        if gs["graph"] is not None:
            gs["graph"].replay()
            self.launches += gs["launches"]
            return gs["loss"]
        return self._step_body_dev(gs)

    def train_step(self, notes, chosen, beat, style, target, seed: int, allreduce=None, world: int = 1):
        """One `fit` batch: forward + primary_loss + backward + gradient exchange + Nadam.  The exchange is either
        `allreduce` (NCCL sum of the flat gradient) followed by the Nadam kernel, or -- when a parallel.PeerNadam is
        attached -- one kernel that does both over peer memory.  Returns the device scalar loss (no sync).
        With `self.graph` (default) and no NCCL collective in the step, the whole step is one CUDA graph launch."""
        if self.graph and allreduce is None and self.profile is None:
            return self._train_step_graph((notes, chosen, beat, style, target), seed, world)
        self.forward(notes, chosen, beat, style, target=target, train=True, seed=seed)
        loss = self.backward()
        if self.peer is not None:
            self.peer.step(self, self._nadam_scalars(), _stream())
            return loss
        if allreduce is not None:
            allreduce(self.gflat)
        self.nadam_step(1.0 / world)
        return loss
