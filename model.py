"""Drop-in for the reference's model.py: same entry points (`build_models`,
`primary_loss`), same argument meaning, but the three returned models run on
the B200 engine (libdeepj_sm100.so) instead of a Keras/TensorFlow graph."""
from constants import *
import music_generator_b200  # noqa: F401  (registers the hyphenated package)
from music_generator_b200.config import ModelConfig
from music_generator_b200.engine import Engine
from music_generator_b200.keras_like import NoteModel, TimeModel, TrainModel
from music_generator_b200.keras_like import primary_loss  # noqa: F401  (model.py:14-20)


def build_models(time_steps=SEQ_LEN, input_dropout=0.2, dropout=0.5, precision='mixed', seed=0,
                 recurrent_activation='hard_sigmoid'):
    """reference model.py:128-169 -> (model, time_model, note_model) sharing one
    set of weights.  `time_steps` is taken from the arrays at call time; it is
    kept for signature compatibility."""
    if TIME_AXIS_LAYERS != 2 or NOTE_AXIS_LAYERS != 2:
        raise ValueError('the B200 engine implements the 2+2 layer DeepJ model of constants.py')
    cfg = ModelConfig(num_styles=NUM_STYLES, num_notes=NUM_NOTES, note_units=NOTE_UNITS,
                      notes_per_bar=NOTES_PER_BAR, seq_len=time_steps, octave_units=OCTAVE_UNITS,
                      style_units=STYLE_UNITS, time_axis_units=TIME_AXIS_UNITS, note_axis_units=NOTE_AXIS_UNITS)
    eng = Engine(cfg, precision=precision, recurrent_activation=recurrent_activation,
                 input_dropout=input_dropout, dropout=dropout)
    eng.init_params(seed)
    return TrainModel(eng, 'model'), TimeModel(eng, 'time_model'), NoteModel(eng, 'note_model')
