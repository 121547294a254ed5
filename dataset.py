"""Data-representation helpers used by the hot path (reference dataset.py:14-26,
84-88) plus the synthetic piano-roll generator that stands in for
`load_all` (dataset.py:39-76 needs a MIDI corpus that the reference does not
ship and the python-midi package)."""
import numpy as np

from constants import *
from util import one_hot


def compute_beat(beat, notes_in_bar):
    return one_hot(beat % notes_in_bar, notes_in_bar)


def compute_genre(genre_id):
    """Uniform mixture over the composers of one genre (dataset.py:20-26)."""
    genre_hot = np.zeros((NUM_STYLES,))
    start = sum(len(s) for s in styles[:genre_id])
    n = len(styles[genre_id])
    genre_hot[start:start + n] = 1 / n
    return genre_hot


def unclamp_midi(sequence):
    """Pad MIN_NOTE silent pitches below the 48-note range (dataset.py:84-88)."""
    return np.pad(sequence, ((0, 0), (MIN_NOTE, 0), (0, 0)), 'constant')


def clamp_midi(sequence):
    return sequence[:, MIN_NOTE:MAX_NOTE, :]


def stagger(data, time_steps):
    """Windows of `time_steps` every NOTES_PER_BAR steps over the sequence left-padded with
    `time_steps` silent frames; labels are the same windows one step later (dataset.py:28-37)."""
    data = list(data)
    padded = [np.zeros_like(data[0])] * time_steps + data
    starts = range(0, len(padded) - time_steps, NOTES_PER_BAR)
    return [padded[i:i + time_steps] for i in starts], [padded[i + 1:i + time_steps + 1] for i in starts]


def load_all(styles, batch_size, time_steps):
    """MIDI corpus -> ([notes, target, beat, style], [target]) training arrays (dataset.py:39-76).
    One style id per composer directory; files shorter than a window are skipped."""
    from concurrent.futures import ThreadPoolExecutor
    from midi_util import load_midi
    from util import get_all_files
    notes, targets, beats, style_rows = [], [], [], []
    composers = [d for group in styles for d in group]
    for style_id, directory in enumerate(composers):
        hot = one_hot(style_id, NUM_STYLES)
        with ThreadPoolExecutor() as pool:           # the reference decodes files on a thread pool, in order
            seqs = list(pool.map(load_midi, get_all_files([directory])))
        for seq in seqs:
            if len(seq) < time_steps:
                continue
            seq = clamp_midi(seq)
            x, y = stagger(seq, time_steps)
            notes += x
            targets += y
            beats += stagger([compute_beat(i, NOTES_PER_BAR) for i in range(len(seq))], time_steps)[0]
            style_rows += stagger([hot] * len(seq), time_steps)[0]
    notes, targets = np.array(notes), np.array(targets)
    return [notes, targets, np.array(beats), np.array(style_rows)], [targets]


def synthetic_all(num_seqs, time_steps=SEQ_LEN, seed=1234):
    """Synthetic stand-in with the shapes/semantics of load_all's return value
    (dataset.py:72-76): ([notes, target, beat, style], [target])."""
    rs = np.random.RandomState(seed)
    roll = np.zeros((num_seqs, time_steps + 1, NUM_NOTES, NOTE_UNITS), dtype=np.float32)
    play = rs.random_sample(roll.shape[:3]) < 0.05
    roll[..., 0] = play
    roll[..., 1] = (rs.random_sample(roll.shape[:3]) < 0.1) * play
    roll[..., 2] = rs.uniform(0.2, 0.8, roll.shape[:3]) * play
    notes, target = roll[:, :-1], roll[:, 1:]
    phase = rs.randint(0, NOTES_PER_BAR, num_seqs)
    beat = np.zeros((num_seqs, time_steps, NOTES_PER_BAR), dtype=np.float32)
    tt = (np.arange(time_steps)[None, :] + phase[:, None]) % NOTES_PER_BAR
    beat[np.arange(num_seqs)[:, None], np.arange(time_steps)[None, :], tt] = 1
    style = np.zeros((num_seqs, time_steps, NUM_STYLES), dtype=np.float32)
    style[np.arange(num_seqs), :, rs.randint(0, NUM_STYLES, num_seqs)] = 1
    return [notes, target, beat, style], [target]
