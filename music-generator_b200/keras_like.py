"""The three model objects `build_models` returns (reference model.py:128-169),
with the slice of the Keras `Model` API the reference's callers use:
`fit` (train.py:29), `predict` (generate.py:108,114), `load_weights` /
`save_weights` (util.py:19, train.py:23), `summary` (util.py:16) and
`get_layer('style')` (visualize.py:13).  NumPy in, NumPy out; all arithmetic
runs in libdeepj_sm100.so through Engine.
"""
from __future__ import annotations

import os
import re
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from .engine import Engine, N, Workspace
from ._lib import NO_DROPOUT

PREDICT_BATCH = 32   # Keras Model.predict default batch_size; it scopes the pitch_bins scramble


def _dev(eng: Engine, a) -> torch.Tensor:
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).to(eng.dev, non_blocking=True)


# Keras auto-names of the reference's weighted layers, in model.py creation order -> our tensor prefixes.
# TimeDistributed wrappers own the variables of the layer they wrap (`time_distributed_4/kernel:0` is lstm_1's kernel).
# Used for WRITING; reading identifies layers by their tensor shapes, which are unique in this model, so a file whose
# auto-names are shifted (a second build_models() in the same Keras session) still loads.
KERAS_LAYERS = [("style", "style", ("W", "b")), ("time_distributed_1", "conv", ("W", "b")),
                ("dense_1", "time0.sd", ("W", "b")), ("time_distributed_4", "time0.lstm", ("W", "U", "b")),
                ("dense_2", "time1.sd", ("W", "b")), ("time_distributed_6", "time1.lstm", ("W", "U", "b")),
                ("dense_3", "note0.sd", ("W", "b")), ("time_distributed_8", "note0.lstm", ("W", "U", "b")),
                ("dense_4", "note1.sd", ("W", "b")), ("time_distributed_10", "note1.lstm", ("W", "U", "b")),
                ("note_dense", "note_dense", ("W", "b")), ("volume_dense", "volume_dense", ("W", "b"))]
_KERAS_VAR = {"W": "kernel:0", "U": "recurrent_kernel:0", "b": "bias:0"}


def keras_h5_tree(state) -> dict:
    """The 28 tensors as the tree `keras.Model.save_weights` writes (Keras 2.0-2.1, TensorFlow backend): root
    attributes `layer_names`, `backend`, `keras_version`; one group per layer with `weight_names` and the
    variables below it under their full TensorFlow names (`dense_1/kernel:0` -> group dense_1, dataset kernel:0)."""
    tree = {"@attrs": {"layer_names": [k for k, _, _ in KERAS_LAYERS], "backend": b"tensorflow",
                       "keras_version": b"2.0.8"}}
    for layer, prefix, parts in KERAS_LAYERS:
        names = [f"{layer}/{_KERAS_VAR[p]}" for p in parts]
        tree[layer] = {"@attrs": {"weight_names": names},
                       layer: {_KERAS_VAR[p]: np.asarray(state[f"{prefix}.{p}"], dtype=np.float32) for p in parts}}
    return tree


def read_keras_h5(f, shapes):
    """Weights from a Keras HDF5 file: `save_weights` layout (layer groups at the root) or a full `model.save` /
    `ModelCheckpoint` file (the same below `model_weights`), as keras.engine.topology.load_weights_from_hdf5_group
    reads them.  Each layer group's variables are taken in `weight_names` order; the layer is identified by the shapes
    of its variables -- (kernel[, recurrent_kernel], bias) -- which are distinct for all 12 weighted layers of
    model.py, with the numeric suffix of the auto-name as the tie-break.  Also reads this package's round-1 flat
    layout (one dataset per tensor name)."""
    from . import h5lite
    if all(k in f for k in shapes):
        return {k: f[k].read() for k in shapes}
    root = f["model_weights"] if ("layer_names" not in f.attrs and "model_weights" in f) else f
    layers = []
    for lname in root.keys():
        g = root[lname]
        if not isinstance(g, h5lite.Group):
            continue
        found = {}
        g.visititems(lambda n, o: found.__setitem__(n[len(g.name) + 1:], o) if isinstance(o, h5lite.Dataset) else None)
        if not found:
            continue
        wn = g.attrs.get("weight_names")
        if wn is not None and len(np.atleast_1d(wn)) == len(found):
            order = [w.decode("utf8") if isinstance(w, bytes) else str(w) for w in np.atleast_1d(wn)]
        else:
            rank = lambda n: (0 if "recurrent" not in n and "kernel" in n else 1 if "recurrent" in n else 2, n)
            order = sorted(found, key=rank)
        if any(o not in found for o in order):
            raise KeyError(f"layer {lname}: weight_names {order} do not match the datasets {sorted(found)}")
        m = re.search(r"_(\d+)$", lname)
        layers.append((int(m.group(1)) if m else 0, lname, [found[o] for o in order]))
    state, used = {}, set()
    for _, prefix, parts in KERAS_LAYERS:
        want = [tuple(shapes[f"{prefix}.{p}"]) for p in parts]
        hits = [L for L in layers if [tuple(d.shape) for d in L[2]] == want and L[1] not in used]
        named = [L for L in hits if L[1] == prefix]            # style / note_dense / volume_dense carry explicit names
        if named:
            hits = named
        if not hits:
            raise KeyError(f"no layer with variables of shapes {want} (for {prefix}) in the HDF5 file")
        L = sorted(hits)[0]
        used.add(L[1])
        for p, d in zip(parts, L[2]):
            state[f"{prefix}.{p}"] = d.read().astype(np.float32)
    return state


class History:
    def __init__(self):
        self.history = {"loss": []}


class _Base:
    def __init__(self, eng: Engine, name: str):
        self.engine, self.name = eng, name

    # -- weights.  The reference saves Keras HDF5 (`out/model.h5`, train.py:23 / util.py:19).  A path ending in .h5 is
    # written as a Keras `save_weights` file and read as one (or as a full `model.save` file) through h5lite, this
    # package's own HDF5 subset -- h5py is not needed.  Any other extension is an .npz keyed by tensor name.
    def save_weights(self, path: str) -> None:
        from . import h5lite
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        state = self.engine.get_params()
        if path.endswith((".h5", ".hdf5")):
            h5lite.write(path, keras_h5_tree(state))
            return
        with open(path, "wb") as f:
            np.savez(f, **state)

    def load_weights(self, path: str) -> None:
        from . import h5lite
        if path.endswith((".h5", ".hdf5")):
            legacy = path[:path.rfind(".")] + ".npz"       # round 1 wrote .npz beside the .h5 name
            if not os.path.exists(path) and os.path.exists(legacy):
                path = legacy
            else:
                with h5lite.File(path) as f:
                    self.engine.set_params(read_keras_h5(f, self.engine.shapes))
                return
        with np.load(path) as z:
            self.engine.set_params({k: z[k] for k in self.engine.shapes})

    def get_weights(self) -> List[np.ndarray]:
        return list(self.engine.get_params().values())

    def count_params(self) -> int:
        return self.engine.num_params

    def summary(self) -> None:
        print(f'Model "{self.name}" (DeepJ biaxial LSTM, B200 engine, {self.engine.precision} gate GEMMs)')
        print("-" * 64)
        for k, s in self.engine.shapes.items():
            print(f"{k:24s} {str(tuple(s)):>20s} {int(np.prod(s)):>12,d}")
        print("-" * 64)
        print(f"Total params: {self.engine.num_params:,d}")

    def get_layer(self, name: str):
        if name != "style":
            raise ValueError(f"No such layer: {name}")
        eng = self.engine

        class _Style:
            """The shared `Dense(STYLE_UNITS, name='style')` layer (model.py:141): weights, and a call that
            embeds style vectors on the device (what visualize.py:16-23 does through a TF session)."""

            def get_weights(self_inner):
                p = eng.get_params()
                return [p["style.W"], p["style.b"]]

            def __call__(self_inner, style_vectors):
                sv = np.ascontiguousarray(style_vectors, dtype=np.float32)
                n, ns = sv.shape
                ws = eng.workspace(n, 1, "fp32", False)
                eng._style(ws, _dev(eng, sv), ns, 0, n, 1)     # dj_style_fwd: emb = style.W + b (linear)
                torch.cuda.synchronize()
                return ws.emb[:n].cpu().numpy().copy()
        return _Style()


class TrainModel(_Base):
    """`model` of model.py:151-152 (inputs [notes, chosen, beat, style])."""

    def predict(self, x: Sequence[np.ndarray], batch_size: int = PREDICT_BATCH) -> np.ndarray:
        notes, chosen, beat, style = x
        outs = []
        for s in range(0, len(notes), batch_size):
            sl = slice(s, s + batch_size)
            ws = self.engine.forward(_dev(self.engine, notes[sl]), _dev(self.engine, chosen[sl]),
                                     _dev(self.engine, beat[sl]), _dev(self.engine, style[sl]),
                                     precision="fp32")
            b, t = notes[sl].shape[0], notes[sl].shape[1]
            outs.append(ws.probs.view(b, t, N, 3).cpu().numpy())
        return np.concatenate(outs, 0)

    def train_on_batch(self, x, y, seed: int = 0, allreduce=None, world: int = 1) -> float:
        e = self.engine
        notes, chosen, beat, style = [_dev(e, a) for a in x]
        target = _dev(e, y[0] if isinstance(y, (list, tuple)) else y)
        return float(e.train_step(notes, chosen, beat, style, target, seed, allreduce, world).item())

    def fit(self, x, y, epochs: int = 1, callbacks: Optional[list] = None, batch_size: int = 16,
            shuffle: bool = True, seed: int = 0, verbose: int = 1, allreduce=None, world: int = 1) -> History:
        """Keras-style epoch loop: shuffled mini-batches (the last one may be short; the pitch_bins scramble is
        scoped to each batch like in Keras).

        Input pipeline: batch i+1 is gathered into pinned host memory and copied to the device on a copy stream
        while batch i trains; the loss is accumulated on the device and read back once per epoch, so there is no
        host synchronisation inside an epoch (the callbacks only need the epoch loss, train.py:22-26)."""
        e = self.engine
        hist = History()
        y0 = y[0] if isinstance(y, (list, tuple)) else y
        arrays = [np.asarray(a) for a in x] + [np.asarray(y0)]
        n = len(arrays[0])
        rs = np.random.RandomState(seed)
        callbacks = callbacks or []
        for cb in callbacks:
            cb.set_model(self)
        bsz = min(batch_size, n)
        pinned = [[torch.empty((bsz,) + a.shape[1:], dtype=torch.float32).pin_memory() for a in arrays] for _ in range(2)]
        devb = [[torch.empty((bsz,) + a.shape[1:], dtype=torch.float32, device=e.dev) for a in arrays] for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=e.dev)
        copied = [None, None]          # event: H2D of this slot finished (the pinned buffers may be rewritten)
        consumed = [None, None]        # event: the train step that read this slot's device buffers finished

        def stage(slot, idx):
            if copied[slot] is not None:
                copied[slot].synchronize()
            b = len(idx)
            for k, a in enumerate(arrays):
                pinned[slot][k].numpy()[:b] = a[idx]
            with torch.cuda.stream(copy_stream):
                if consumed[slot] is not None:
                    copy_stream.wait_event(consumed[slot])
                for k in range(len(arrays)):
                    devb[slot][k][:b].copy_(pinned[slot][k][:b], non_blocking=True)
                copied[slot] = torch.cuda.Event()
                copied[slot].record(copy_stream)
            return b

        step = 0
        main = torch.cuda.current_stream(e.dev)
        for ep in range(epochs):
            order = rs.permutation(n) if shuffle else np.arange(n)
            starts = list(range(0, n, batch_size))
            tot = torch.zeros((), dtype=torch.float64, device=e.dev)
            nb = stage(0, order[starts[0]:starts[0] + batch_size])
            for bi, s0 in enumerate(starts):
                slot, b = bi & 1, nb
                main.wait_event(copied[slot])
                if bi + 1 < len(starts):
                    s1 = starts[bi + 1]
                    nb = stage(slot ^ 1, order[s1:s1 + batch_size])
                d = [t[:b] for t in devb[slot]]
                loss = e.train_step(d[0], d[1], d[2], d[3], d[4], seed * 1000003 + step, allreduce, world)
                tot += loss.double().reshape(()) * b
                consumed[slot] = torch.cuda.Event()
                consumed[slot].record(main)
                step += 1
            if world > 1:
                # every rank must see the same epoch loss, or rank-local callbacks (early stopping) would make the
                # ranks leave the epoch loop at different times and the next gradient exchange would wait forever
                # The same all-reduce carries the health of the fused gradient exchange: if any rank's bounded wait
                # gave up during this epoch, every rank stops HERE, before ModelCheckpoint could save anything.
                from . import parallel
                bad = e.peer.status_word().double().reshape(()) if e.peer is not None else torch.zeros_like(tot)
                tot_n = parallel.sum_over_ranks(torch.stack([tot, torch.full_like(tot, float(n)), bad]))
                if float(tot_n[2].item()) != 0.0:
                    raise RuntimeError("the peer-memory gradient exchange timed out on at least one rank during this epoch; "
                                       "no update was applied after that step and nothing was saved")
                logs = {"loss": float(tot_n[0].item()) / max(float(tot_n[1].item()), 1.0)}
            else:
                logs = {"loss": float(tot.item()) / max(n, 1)}
            hist.history["loss"].append(logs["loss"])
            if verbose:
                print(f"Epoch {ep + 1}/{epochs} - loss: {logs['loss']:.4f}")
            stop = False
            for cb in callbacks:
                cb.on_epoch_end(ep, logs)
                stop = stop or getattr(cb, "stop_training", False)
            if stop:
                break
        return hist


class TimeModel(_Base):
    """`time_model` of model.py:155: [notes, beat, style] -> time_out [G,T,48,Ut]."""

    def predict(self, x: Sequence[np.ndarray], batch_size: int = PREDICT_BATCH) -> np.ndarray:
        e = self.engine
        notes, beat, style = x
        outs = []
        d = {s: NO_DROPOUT for s in range(1, 13)}
        for s in range(0, len(notes), batch_size):
            n_, b_, s_ = [_dev(e, a[s:s + batch_size]) for a in (notes, beat, style)]
            B, T = n_.shape[0], n_.shape[1]
            ws = e.workspace(B, T, "fp32", False)
            e.forward_time(ws, n_, T * N * 3, b_, T * 16, B, T, d, False, False, style=s_,
                           style_bstride=T * e.cfg.num_styles, style_tstride=e.cfg.num_styles)
            outs.append(ws.h[1].view(B, T, N, -1).cpu().numpy())
        return np.concatenate(outs, 0)


class NoteModel(_Base):
    """`note_model` of model.py:157-167: [note_features [G,1,48,Ut], chosen [G,1,48,3],
    style [G,1,23]] -> [G,1,48,3]."""

    def predict(self, x: Sequence[np.ndarray], batch_size: int = PREDICT_BATCH) -> np.ndarray:
        e = self.engine
        feats, chosen, style = x
        outs = []
        d = {s: NO_DROPOUT for s in range(1, 13)}
        for s in range(0, len(feats), batch_size):
            f_, c_, s_ = [_dev(e, a[s:s + batch_size]) for a in (feats, chosen, style)]
            B, T = f_.shape[0], f_.shape[1]
            ws = e.workspace(B, T, "fp32", False)
            e._style(ws, s_, T * e.cfg.num_styles, e.cfg.num_styles, B, T)
            e.forward_note(ws, f_.view(B * T * N, -1), 0, T * N, c_, T * N * 3, B, T, d, False, False)
            outs.append(ws.probs.view(B, T, N, 3).cpu().numpy())
        return np.concatenate(outs, 0)


def primary_loss(y_true: np.ndarray, y_pred: np.ndarray) -> np.ndarray:
    """Host restatement of model.py:14-20 for callers that want the per-(b,t)
    loss map of already-computed predictions (the training path computes the
    loss on device in dj_head_loss)."""
    y_true = np.asarray(y_true, dtype=np.float32); y_pred = np.asarray(y_pred, dtype=np.float32)
    eps = np.float32(1e-7)

    def bce(t, o):
        o = np.clip(o, eps, 1 - eps)
        return -(t * np.log(o) + (1 - t) * np.log1p(-o)).mean(-1)
    played = y_true[..., 0]
    out = bce(y_true[..., 0], y_pred[..., 0])
    out = out + bce(y_true[..., 1], played * y_pred[..., 1] + (1 - played) * y_true[..., 1])
    d = y_true[..., 2] - (played * y_pred[..., 2] + (1 - played) * y_true[..., 2])
    return out + (d * d).mean(-1)
