"""Dumps the style embedding of every composer for http://projector.tensorflow.org/ (reference visualize.py:11-43):
`out/style_embedding_vec.tsv` (one 64-d row per style, the `style` Dense layer applied to the identity) and
`out/style_embedding_labels.tsv` (Genre / Artist).  The layer runs on the device through the engine
(dj_style_fwd) instead of a TensorFlow session."""
import os

import numpy as np

from constants import *
from util import *


def style_labels():
    """[1 + NUM_STYLES, 2] table: header row, then (genre, artist path) per style (visualize.py:27-40)."""
    labels = [[g] * len(styles[i]) for i, g in enumerate(genre)]
    labels = [y for x in labels for y in x]
    styles_labels = [y for x in styles for y in x]
    table = np.hstack([np.reshape(labels, [-1, 1]), np.reshape(styles_labels, [-1, 1])])
    return np.vstack([['Genre', 'Artist'], table])


def main():
    models = build_or_load()
    style_layer = models[0].get_layer('style')
    print('Creating input')
    all_styles = np.identity(NUM_STYLES)          # all possible styles
    embedding = style_layer(all_styles)
    print('Writing to out directory')
    os.makedirs(OUT_DIR, exist_ok=True)
    np.savetxt(os.path.join(OUT_DIR, 'style_embedding_vec.tsv'), embedding, delimiter='\t')
    np.savetxt(os.path.join(OUT_DIR, 'style_embedding_labels.tsv'), style_labels(), delimiter='\t', fmt='%s')


if __name__ == '__main__':
    main()
