"""CPU study (oracle with rounding injected): which bf16 operand roundings of the training path fit the
north_star tolerance (probabilities / loss within 1e-3 relative of the fp64 model)?

    python tools/precision_study.py [B] [T]

Every variant runs the fp64 oracle forward with dropout masks on, rounding the named operands of the LSTM
layers to bf16 (`1`), to a hi+lo pair of bf16 (`2`, what a two-pass split MMA sees), or not at all (`0`):
  A = gate-GEMM input rows, W = gate-GEMM kernel, H = recurrent input h_{t-1}, U = recurrent kernel.
Test infrastructure: imports oracle/, never imported by the product path.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import deepj_oracle as O   # noqa: E402
import helpers                          # noqa: E402


def rnd(x, mode):
    if mode == 0:
        return x
    t16 = torch.bfloat16 if mode in (1, 2) else torch.float16      # 1/2: bf16, bf16 hi+lo; 3/4: fp16, fp16 hi+lo
    hi = x.to(t16).to(x.dtype)
    if mode in (1, 3):
        return hi
    lo = (x - hi).to(t16).to(x.dtype)
    return hi + lo


def make_lstm(a, w, h_, u_):
    def lstm_seq(x, W, U, b, recurrent_activation="hard_sigmoid", return_state=False):
        S, steps, _ = x.shape
        u = U.shape[0]
        act = O.hard_sigmoid if recurrent_activation == "hard_sigmoid" else torch.sigmoid
        h = x.new_zeros(S, u); c = x.new_zeros(S, u)
        zxs = (rnd(x, a) @ rnd(W, w) + b).float().double().unbind(1)     # Z is stored in fp32
        Ur = rnd(U, u_)
        hs = []
        for t in range(steps):
            z = zxs[t] + rnd(h, h_) @ Ur
            i = act(z[:, :u]); f = act(z[:, u:2 * u]); g = torch.tanh(z[:, 2 * u:3 * u]); o = act(z[:, 3 * u:])
            c = f * c + i * g
            h = o * torch.tanh(c)
            hs.append(h)
        return torch.stack(hs, dim=1)
    return lstm_seq


VARIANTS = [(1, 1, 1, 1), (1, 1, 0, 0), (0, 0, 1, 1), (2, 2, 1, 1), (2, 2, 2, 2), (1, 0, 0, 0), (0, 1, 0, 0),
            (0, 0, 1, 0), (0, 0, 0, 1), (3, 3, 3, 3), (3, 3, 0, 0), (0, 0, 3, 3), (4, 4, 3, 3), (4, 4, 4, 3), (4, 4, 4, 4)]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    cfg = O.Config()
    p = O.init_params(cfg, 0, torch.float64)
    batch = [t.double() for t in O.synthetic_batch(cfg, B, T, 1234)]
    masks = helpers.oracle_masks(cfg, B, T, 7)
    orig = O.lstm_seq
    with torch.no_grad():
        ref = O.model_forward(p, cfg, *batch[:4], masks=masks)
        lref = float(O.primary_loss(batch[4], ref))
        print(f"B={B} T={T}  ref loss {lref:.6f}  |p|max per channel {ref.abs().amax((0, 1, 2)).tolist()}")
        print("variant(A W H U)   loss_rel   play_abs  replay_abs  vol_abs   vol_abs/max|vol|   elem_rel(floor .05 / .1 / .25)")
        for v in VARIANTS:
            O.lstm_seq = make_lstm(*v)
            out = O.model_forward(p, cfg, *batch[:4], masks=masks)
            O.lstm_seq = orig
            l = float(O.primary_loss(batch[4], out))
            e = (out - ref).abs()
            ch = e.amax((0, 1, 2)).tolist()
            rel = [(e / ref.abs().clamp(min=fl)).max().item() for fl in (0.05, 0.1, 0.25)]
            print(f"{v}   {abs(l - lref) / lref:.2e}   {ch[0]:.2e}  {ch[1]:.2e}  {ch[2]:.2e}   "
                  f"{ch[2] / ref[..., 2].abs().max().item():.2e}   {rel[0]:.2e} {rel[1]:.2e} {rel[2]:.2e}", flush=True)


if __name__ == "__main__":
    main()
