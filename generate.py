"""Drop-in for the reference's generate.py: `python generate.py [--bars N]
[--styles i j ...]` and `generate(models, num_bars, styles)` (a generator that
yields, per timestep, a list of one [48,3] array per sequence,
generate.py:98-121).  The autoregressive loop itself -- window recompute,
note-by-note LSTM, temperature, Bernoulli draws -- runs on the GPU
(music_generator_b200.sampler); the host only supplies the uniform stream that
np.random.random() would have produced, in the same order."""
import argparse
import os

import numpy as np

from constants import *
from dataset import compute_genre, unclamp_midi
from util import build_or_load, one_hot
from midi_util import midi_encode
from music_generator_b200 import smf as midi
from music_generator_b200.sampler import generate_events


def apply_temperature(prob, temperature):
    """generate.py:81-91 (host copy for API parity; the device sampler applies
    the same float32 transform)."""
    if temperature != 1:
        x = -np.log(1 / prob - 1)
        prob = 1 / (1 + np.exp(-x / temperature))
    return prob


def generate(models, num_bars, styles, uniforms=None, default_temp=1):
    print('Generating with styles:', styles)
    eng = models[1].engine
    steps = NOTES_PER_BAR * num_bars
    if uniforms is None:
        # worst case 2 draws per (timestep, sequence, note): same values, same order
        # as the reference's successive np.random.random() calls
        uniforms = np.random.random_sample(2 * steps * len(styles) * NUM_NOTES)
    events, _ = generate_events(eng, styles, steps, uniforms, stream_mode=0, default_temp=default_temp)
    for t in range(steps):
        yield [events[t, i] for i in range(len(styles))]


def generate_batch(models, num_bars, styles, seed=0, default_temp=1):
    """Many independent style-conditioned sequences (BASELINE configs[3]): every (timestep, sequence,
    note) owns its two uniforms U[t,g,n,:] -- drawn from RandomState(seed) for the WHOLE batch -- so the
    result does not depend on how the sequences are sharded.  Under torchrun rank r generates the 32-sequence
    predict chunks r, r+world, ... (whole chunks keep the pitch-bins scope of a single-GPU run, model.py:43-49);
    no collective is involved.  Returns (indices, events[steps, len(indices), 48, 3])."""
    from music_generator_b200 import parallel
    rank, world, local = parallel.env_world()
    eng = models[1].engine
    steps = NOTES_PER_BAR * num_bars
    G = len(styles)
    u = np.random.RandomState(seed).random_sample((steps, G, NUM_NOTES, 2))
    mine = parallel.shard_sequence_chunks(G, rank, world)
    if not mine:
        return mine, np.zeros((steps, 0, NUM_NOTES, NOTE_UNITS), dtype=np.float32)
    events, _ = generate_events(eng, [styles[g] for g in mine], steps, np.ascontiguousarray(u[:, mine]),
                                stream_mode=1, default_temp=default_temp)
    return mine, events


def batch_styles(G, seed=0):
    """Styles of a batched run: the three genre mixtures of dataset.compute_genre cycled with random 3-composer
    mixtures."""
    rs = np.random.RandomState(seed)
    out = []
    for g in range(G):
        if g % 2 == 0:
            out.append(compute_genre((g // 2) % len(genre)))
        else:
            out.append(np.mean([one_hot(i, NUM_STYLES) for i in rs.choice(NUM_STYLES, 3, replace=False)], axis=0))
    return out


def write_file(name, results):
    """generate.py:123-134: one .mid per generated sequence (unclamp to the 128-pitch
    roll, midi_encode, write a Standard MIDI File)."""
    results = zip(*list(results))
    for i, result in enumerate(results):
        fpath = os.path.join(SAMPLES_DIR, name + '_' + str(i) + '.mid')
        print('Writing file', fpath)
        os.makedirs(os.path.dirname(fpath), exist_ok=True)
        midi.write_midifile(fpath, midi_encode(unclamp_midi(np.array(result))))


def main():
    parser = argparse.ArgumentParser(description='Generates music.')
    parser.add_argument('--bars', default=32, type=int, help='Number of bars to generate')
    parser.add_argument('--styles', default=None, type=int, nargs='+', help='Styles to mix together')
    parser.add_argument('--batch', default=0, type=int,
                        help='Generate this many independent sequences (sharded over ranks under torchrun)')
    parser.add_argument('--seed', default=0, type=int, help='Seed of the indexed uniform stream of --batch')
    args = parser.parse_args()
    if args.batch:
        # BASELINE configs[3]: N independent style-conditioned sequences; under torchrun every rank generates its
        # own share (whole 32-sequence predict chunks, no collective) and writes its own, globally numbered files
        import torch
        from music_generator_b200 import parallel
        rank, world, local = parallel.env_world()
        torch.cuda.set_device(local)
        models = build_or_load()
        styles = batch_styles(args.batch, args.seed)
        mine, events = generate_batch(models, args.bars, styles, seed=args.seed)
        for j, g in enumerate(mine):
            fpath = os.path.join(SAMPLES_DIR, 'batch_' + str(g) + '.mid')
            os.makedirs(os.path.dirname(fpath), exist_ok=True)
            midi.write_midifile(fpath, midi_encode(unclamp_midi(events[:, j])))
        print('rank', rank, 'wrote', len(mine), 'of', args.batch, 'sequences to', SAMPLES_DIR)
        return

    from music_generator_b200 import parallel
    if parallel.env_world()[0] != 0:
        return                      # the 1..3 sequence path is one GPU's work: only rank 0 generates and writes
    models = build_or_load()
    styles = [compute_genre(i) for i in range(len(genre))]
    if args.styles:
        styles = [np.mean([one_hot(i, NUM_STYLES) for i in args.styles], axis=0)]
    write_file('output', generate(models, args.bars, styles))


if __name__ == '__main__':
    main()
