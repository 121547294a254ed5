"""Piano-roll <-> MIDI codec: the data format on the far side of the hot path
(reference midi_util.py:9-210).  Re-implemented as a
change-point encoder and a streaming window decoder over `music_generator_b200.smf`
(python-midi is not installable here); behaviour is pinned by the reference's own
test.py cases, restated in tests/test_midi.py."""
import os

import numpy as np

from constants import *
import music_generator_b200  # noqa: F401
from music_generator_b200 import smf as midi


def midi_encode(note_seq, resolution=NOTES_PER_BEAT, step=1):
    """Piano roll [T, P, 3] (play, replay, volume) -> one-track Pattern.

    Events are emitted only on timesteps whose play vector differs from the previous
    one (so a replay flag on an otherwise unchanged frame is dropped -- reference
    midi_util.py:35,56); within a timestep pitches are visited in ascending order;
    delta ticks count from the last emitted event; the end-of-track delta is the
    number of unchanged frames since the last change (midi_util.py:71-92)."""
    note_seq = np.asarray(note_seq)
    play, replay, volume = note_seq[:, :, 0], note_seq[:, :, 1], note_seq[:, :, 2]
    pattern = midi.Pattern(resolution=resolution)
    track = midi.Track()
    pattern.append(track)
    T = len(play)
    prev = np.zeros_like(play[0])
    last_tick, idle = 0, 0
    for t in range(T):
        cur = play[t]
        if np.array_equal(prev, cur):
            idle += 1
            continue
        idle = 0
        on_now, on_before = cur > 0, prev > 0
        changed = np.nonzero((on_now != on_before) | (on_now & on_before & (replay[t] > 0)))[0]
        for p in changed:
            delta = (t - last_tick) * step
            # the volume head is linear and unbounded (model.py:95): clamp instead of letting 128 wrap to 0 or a
            # negative value wrap to a loud note in the 7-bit data byte
            vel = min(max(int(volume[t][p] * MAX_VELOCITY), 0), MAX_VELOCITY)
            if on_now[p] and not on_before[p]:
                track.append(midi.NoteOnEvent(tick=delta, velocity=vel, pitch=int(p)))
            elif on_before[p] and not on_now[p]:
                track.append(midi.NoteOffEvent(tick=delta, pitch=int(p)))
            else:   # held and re-articulated
                track.append(midi.NoteOffEvent(tick=delta, pitch=int(p)))
                track.append(midi.NoteOnEvent(tick=0, velocity=vel, pitch=int(p)))
            last_tick = t
        prev = cur
    for p in np.nonzero(prev > 0)[0]:          # release whatever still sounds after the last frame
        track.append(midi.NoteOffEvent(tick=(T - last_tick) * step, pitch=int(p)))
        last_tick, idle = T, 0
    track.append(midi.EndOfTrackEvent(tick=idle))
    return pattern


class _WindowDecoder:
    """One track at tick resolution, folded into windows of `step` ticks.  A window's volume is the
    max over its ticks, its replay flag the OR; the frame that opens the next window is carried over
    (reference midi_util.py:118-141)."""

    def __init__(self, classes, step):
        self.step = step
        self.vol = np.zeros(classes)        # newest tick frame (mutable by events)
        self.rep = np.zeros(classes)
        self.older_max = np.zeros(classes)  # ticks of the open window before the newest
        self.older_rep = np.zeros(classes)
        self.first = None                   # volume of the window's first tick once it is frozen
        self.prev = None                    # frame before the newest (None right after a window closed)
        self.count = 1
        self.volumes, self.replays = [], []

    def advance(self):
        if self.first is None:
            self.first = self.vol.copy()
        self.older_max = np.maximum(self.older_max, self.vol)
        self.older_rep = self.older_rep + self.rep
        self.prev = self.vol.copy()
        self.rep = np.zeros_like(self.rep)
        self.count += 1
        if self.count > self.step:
            self.replays.append(np.minimum(self.older_rep, 1))
            self.volumes.append(self.older_max)
            self.older_max = np.zeros_like(self.vol)
            self.older_rep = np.zeros_like(self.vol)
            self.first, self.prev, self.count = None, None, 1

    def note_on(self, pitch, velocity):
        self.vol[pitch] = velocity / MAX_VELOCITY
        if self.prev is not None and self.prev[pitch] > 0 and self.vol[pitch] > 0:
            self.rep[pitch] = 1                 # struck again while sounding: a replay at the old volume
            self.vol[pitch] = self.prev[pitch]

    def note_off(self, pitch):
        self.vol[pitch] = 0

    def finish(self):
        self.replays.append(np.minimum(self.older_rep + self.rep, 1))
        self.volumes.append(self.vol.copy() if self.first is None else self.first)
        return np.array(self.replays), np.array(self.volumes)


def midi_decode(pattern, classes=MIDI_MAX_NOTES, step=None):
    """Pattern -> piano roll [T, classes, 3]; tracks are summed and clipped to 1."""
    if step is None:
        step = pattern.resolution // NOTES_PER_BEAT
    merged_r = merged_v = None
    for track in pattern:
        dec = _WindowDecoder(classes, step)
        for event in track:
            for _ in range(event.tick):
                dec.advance()
            if isinstance(event, midi.EndOfTrackEvent):
                break
            if isinstance(event, midi.NoteOnEvent):
                dec.note_on(*event.data)
            elif isinstance(event, midi.NoteOffEvent):
                dec.note_off(event.data[0])
        r, v = dec.finish()
        if merged_v is None:
            merged_r, merged_v = r, v
        else:
            n = max(len(v), len(merged_v))
            pad = lambda a: np.pad(a, ((0, n - len(a)), (0, 0)), 'constant')
            merged_r, merged_v = pad(merged_r) + pad(r), pad(merged_v) + pad(v)
    merged = np.stack([np.ceil(merged_v), merged_r, merged_v], axis=2)
    return np.minimum(merged, 1)


def load_midi(fname):
    """Decode with an .npy cache under CACHE_DIR (reference midi_util.py:193-210)."""
    cache_path = os.path.join(CACHE_DIR, fname + '.npy')
    try:
        note_seq = np.load(cache_path)
    except Exception:
        note_seq = midi_decode(midi.read_midifile(fname))
        os.makedirs(os.path.dirname(cache_path), exist_ok=True)
        np.save(cache_path, note_seq)
    assert note_seq.ndim == 3 and note_seq.shape[1] == MIDI_MAX_NOTES and note_seq.shape[2] == 3, note_seq.shape
    assert (note_seq >= 0).all() and (note_seq <= 1).all()
    return note_seq
