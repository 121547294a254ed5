"""Generates tests/golden/deepj_small.npz from the fp64 oracle.

The reference (Keras/TF) cannot run in this container, so these are NOT
reference outputs: they freeze the oracle's fp64 results on seeded inputs so
that (a) oracle regressions are caught on CPU and (b) the CUDA path is compared
against a committed fixture on the GPU box, where /root/reference and this
generator's inputs do not have to be recomputed.  Re-run: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import deepj_oracle as O  # noqa: E402
import helpers  # noqa: E402

B, T, DATA_SEED, W_SEED, MASK_SEED = 2, 4, 1234, 0, 7


def main():
    cfg = O.Config()
    p = O.init_params(cfg, W_SEED, torch.float64)
    batch = O.synthetic_batch(cfg, B, T, DATA_SEED, torch.float64)
    notes, chosen, beat, style, target = batch
    out = {}
    probs = O.model_forward(p, cfg, notes, chosen, beat, style)
    out["predict_probs"] = probs.numpy()
    out["predict_loss"] = O.primary_loss(target, probs).numpy()
    masks = helpers.oracle_masks(cfg, B, T, MASK_SEED)
    loss, tprobs, grads = O.loss_and_grads(p, cfg, notes, chosen, beat, style, target, masks)
    out["train_loss"] = loss.numpy()
    out["train_probs"] = tprobs.numpy()
    for k, g in grads.items():
        g = g.numpy().ravel()
        out[f"grad_sum/{k}"] = np.array([g.sum(), np.abs(g).sum(), np.abs(g).max()])
        out[f"grad_head/{k}"] = g[:16].copy()
    st = O.NadamState()
    p2 = O.nadam_step({k: v.clone() for k, v in p.items()}, grads, st)
    for k in ("style.W", "conv.W", "time0.lstm.U", "note1.lstm.W", "note_dense.W"):
        out[f"nadam_head/{k}"] = p2[k].numpy().ravel()[:16].copy()
    # generation: 2 timesteps, G=1 style mix of generate.py --styles 0 5 12
    sty = np.mean([np.eye(cfg.num_styles)[i] for i in (0, 5, 12)], axis=0)
    u = np.random.RandomState(42).random_sample(2 * 48 * 2)
    ev, info = O.generate(p, cfg, [sty], 2, u, mode="literal", dtype=torch.float64, return_probs=True)
    out["gen_events"] = ev
    out["gen_probs"] = info["probs"]
    out["gen_used"] = np.array(info["uniforms_used"])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "deepj_small.npz"), **out)
    print("wrote", len(out), "arrays; train_loss", float(loss))


if __name__ == "__main__":
    main()
