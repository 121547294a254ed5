"""Shared test helpers: numpy replay of the kernels' dropout hash, oracle glue."""
import numpy as np
import torch

from oracle import deepj_oracle as O

M32 = np.uint64(0xFFFFFFFF)


def mix32(x):
    x = x.astype(np.uint64) & M32
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x7feb352d)) & M32
    x ^= x >> np.uint64(15); x = (x * np.uint64(0x846ca68b)) & M32
    x ^= x >> np.uint64(16)
    return x


def site_key(seed, site):
    k = mix32(np.array([(seed & 0xFFFFFFFF) ^ ((0x9E3779B9 * (site + 1)) & 0xFFFFFFFF)], dtype=np.uint64))
    return mix32((k + np.uint64(seed >> 32)) & M32)[0]


def keep_mask(seed, site, rate, rows, F):
    """numpy twin of dj_keep (csrc/dj_common.cuh): element index = row*roundup4(F)+f."""
    if rate == 0:
        return np.ones((rows, F), dtype=np.float32)
    key = site_key(seed, site)
    thr = np.uint64(int(rate * 4294967296.0))
    ld4 = (F + 3) & ~3
    e = (np.arange(rows, dtype=np.uint64)[:, None] * np.uint64(ld4) + np.arange(F, dtype=np.uint64)[None, :])
    t256 = rate * 256.0
    if t256 == int(t256):
        w = mix32(((e >> np.uint64(2)) + key) & M32)
        byte = (w >> (np.uint64(8) * (e & np.uint64(3)))) & np.uint64(0xFF)
        return (byte >= (thr >> np.uint64(24))).astype(np.float32)
    w = mix32((e + key) & M32)
    return (w >= thr).astype(np.float32)


def oracle_masks(cfg, B, T, seed, input_dropout=0.2, dropout=0.5):
    """All 12 keep masks in the oracle's shapes, replaying the kernels' hash."""
    N = cfg.num_notes
    M = B * T * N
    out = {}
    shapes = O.dropout_site_shapes(cfg, B, T)
    for k, shp in shapes.items():
        if shp is None:
            continue
        s = int(k[1:])
        rate = input_dropout if s <= 3 else dropout
        F = shp[-1]
        rows = int(np.prod(shp[:-1]))
        out[k] = torch.tensor(keep_mask(seed, s, rate, rows, F).reshape(shp))
    return out


def to_oracle_params(state, dtype=torch.float64):
    return {k: torch.tensor(np.asarray(v), dtype=dtype) for k, v in state.items()}


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
