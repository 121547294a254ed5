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
