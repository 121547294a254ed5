"""Drop-in for the reference's train.py: `python train.py` builds or loads the
model and calls `models[0].fit(...)` with best-loss checkpointing and early
stopping (train.py:18-29).  Under torchrun it trains data-parallel: one process
per GPU, each on its shard of the sequences; the gradient exchange and the Nadam
update are one kernel over NVLink peer memory (parallel.make_step_exchange), the
epoch loss the callbacks see is the mean over all ranks."""
import argparse
import os

import numpy as np

from constants import *
from dataset import load_all, synthetic_all
from util import build_or_load


class ModelCheckpoint:
    """keras.callbacks.ModelCheckpoint(monitor='loss', save_best_only=True,
    save_weights_only=True) -- train.py:23."""

    def __init__(self, path):
        self.path, self.best, self.model = path, np.inf, None

    def set_model(self, model):
        self.model = model

    def on_epoch_end(self, epoch, logs):
        if logs['loss'] < self.best:
            self.best = logs['loss']
            self.model.save_weights(self.path)


class EarlyStopping:
    """keras.callbacks.EarlyStopping(monitor='loss', patience=5) -- train.py:24."""

    def __init__(self, patience=5):
        self.patience, self.best, self.wait, self.stop_training = patience, np.inf, 0, False

    def set_model(self, model):
        pass

    def on_epoch_end(self, epoch, logs):
        if logs['loss'] < self.best:
            self.best, self.wait = logs['loss'], 0
        else:
            self.wait += 1
            self.stop_training = self.wait >= self.patience


def train(models, epochs=1000, num_seqs=256):
    print('Loading data')
    # the reference trains on a MIDI corpus under data/ that it does not ship (README.md:20); use it when it is
    # there, otherwise synthetic piano-rolls of the same shape stand in
    if any(os.path.isdir(d) for group in styles for d in group):
        train_data, train_labels = load_all(styles, BATCH_SIZE, SEQ_LEN)
    else:
        train_data, train_labels = synthetic_all(num_seqs, SEQ_LEN)
    from music_generator_b200 import parallel
    rank, world, _ = parallel.init_distributed()
    allreduce, peer = None, None
    if world > 1:
        # equal shards (the remainder of n / world is dropped): every rank must run the same number of steps
        n = len(train_data[0])
        idx = parallel.shard_indices(n, rank, world)[:n // world]
        train_data = [a[idx] for a in train_data]
        train_labels = [a[idx] for a in train_labels]
        allreduce, peer = parallel.make_step_exchange(models[0].engine, world)
    # every rank stops on the same (global) epoch loss; only rank 0 writes the checkpoint
    cbs = ([ModelCheckpoint(MODEL_FILE)] if rank == 0 else []) + [EarlyStopping(patience=5)]
    print('Training')
    models[0].fit(train_data, train_labels, epochs=epochs, callbacks=cbs, batch_size=BATCH_SIZE,
                  allreduce=allreduce, world=world, verbose=1 if rank == 0 else 0)
    if peer is not None:
        peer.raise_if_timed_out()
        peer.close(models[0].engine)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--epochs', type=int, default=1000)
    ap.add_argument('--num-seqs', type=int, default=256)
    args = ap.parse_args()
    from music_generator_b200 import parallel
    parallel.init_distributed()          # under torchrun: selects this rank's GPU before the model is built
    models = build_or_load()
    train(models, args.epochs, args.num_seqs)


if __name__ == '__main__':
    main()
