"""Configuration contract of the hot path -- same names and values as the
reference's constants.py:1-84 so `from constants import *` callers keep working."""
import os

# genre -> composer directories (reference constants.py:4-40)
_CORPUS = {
    'baroque': ['bach', 'handel', 'pachelbel'],
    'classical': ['burgmueller', 'clementi', 'haydn', 'beethoven', 'brahms', 'mozart'],
    'romantic': ['balakirew', 'borodin', 'brahms', 'chopin', 'debussy', 'liszt', 'mendelssohn',
                 'moszkowski', 'mussorgsky', 'rachmaninov', 'schubert', 'schumann', 'tchaikovsky', 'tschai'],
}
genre = list(_CORPUS)
styles = [['data/%s/%s' % (g, c) for c in _CORPUS[g]] for g in genre]
NUM_STYLES = sum(len(s) for s in styles)          # 23

DEFAULT_RES, MIDI_MAX_NOTES, MAX_VELOCITY = 96, 128, 127

NUM_OCTAVES, OCTAVE = 4, 12
MIN_NOTE = 36
MAX_NOTE = MIN_NOTE + NUM_OCTAVES * OCTAVE
NUM_NOTES = MAX_NOTE - MIN_NOTE                   # 48

BEATS_PER_BAR, NOTES_PER_BEAT = 4, 4
NOTES_PER_BAR = NOTES_PER_BEAT * BEATS_PER_BAR    # 16

BATCH_SIZE = 16
SEQ_LEN = 8 * NOTES_PER_BAR                       # 128

OCTAVE_UNITS = 64
STYLE_UNITS = 64
NOTE_UNITS = 3
TIME_AXIS_UNITS = 256
NOTE_AXIS_UNITS = 128
TIME_AXIS_LAYERS = 2
NOTE_AXIS_LAYERS = 2

OUT_DIR = 'out'
MODEL_DIR = os.path.join(OUT_DIR, 'models')
MODEL_FILE = os.path.join(OUT_DIR, 'model.h5')
SAMPLES_DIR = os.path.join(OUT_DIR, 'samples')
CACHE_DIR = os.path.join(OUT_DIR, 'cache')
