"""Host helpers on the entry path (reference util.py:8-33)."""
import os

import numpy as np

from constants import *


def one_hot(i, nb_classes):
    arr = np.zeros((nb_classes,))
    arr[i] = 1
    return arr


def build_or_load(allow_load=True):
    """reference util.py:13-23: build, print the summary, try to load MODEL_FILE;
    any load failure falls back to the freshly initialised weights."""
    from model import build_models
    models = build_models()
    models[0].summary()
    if allow_load:
        try:
            models[0].load_weights(MODEL_FILE)
            print('Loaded model from file.')
        except Exception:
            print('Unable to load model from file.')
    return models


def get_all_files(paths):
    """Every .mid under the given directories (reference util.py:25-33)."""
    found = []
    for path in paths:
        for root, _, files in os.walk(path):
            found += [os.path.join(root, f) for f in sorted(files) if f.endswith('.mid')]
    return [f for f in found if os.path.isfile(f)]
