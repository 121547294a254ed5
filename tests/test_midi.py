"""Known-answer tests of the piano-roll <-> MIDI codec.  The six cases restate the reference's own
test.py (test.py:7-193: same input rolls / event lists, same expected values); the rest covers the
SMF reader/writer that stands in for python-midi."""
import io
import os

import numpy as np
import pytest

import constants as K
from midi_util import midi_decode, midi_encode
from music_generator_b200 import smf as midi

ROLL_PLAY = [[0, 1, 0, 0], [0, 1, 0, 0], [0, 1, 0, 1], [0, 1, 0, 1], [0, 0, 0, 1], [0, 0, 0, 0]]


def roll(play, replay=None, vol=0.5):
    play = np.array(play, dtype=float)
    replay = np.zeros_like(play) if replay is None else np.array(replay, dtype=float)
    return np.stack([play, replay, play * vol], 2)


def one_track(events, resolution=96):
    pat = midi.Pattern(resolution=resolution)
    tr = midi.Track()
    pat.append(tr)
    tr.extend(events)
    return pat


def test_encode_events_and_deltas():                       # test.py:7-53
    pat = midi_encode(roll(ROLL_PLAY), step=1)
    assert pat.resolution == K.NOTES_PER_BEAT and len(pat) == 1
    tr = pat[0]
    assert len(tr) == 5 and isinstance(tr[-1], midi.EndOfTrackEvent)
    on1, on2, off1, off2 = tr[:-1]
    assert [type(e) for e in (on1, on2, off1, off2)] == [midi.NoteOnEvent, midi.NoteOnEvent, midi.NoteOffEvent, midi.NoteOffEvent]
    assert (on1.tick, on1.pitch) == (0, 1) and (on2.tick, on2.pitch) == (2, 3)
    assert (off1.tick, off1.pitch) == (2, 1) and (off2.tick, off2.pitch) == (1, 3)
    assert on1.velocity == int(0.5 * 127)


def test_decode_downsamples_to_step():                     # test.py:55-77
    pat = one_track([midi.NoteOnEvent(tick=0, velocity=127, pitch=0), midi.NoteOnEvent(tick=96, velocity=127, pitch=1),
                     midi.NoteOffEvent(tick=0, velocity=127, pitch=0), midi.NoteOffEvent(tick=48, velocity=127, pitch=1),
                     midi.EndOfTrackEvent(tick=1)])
    seq = midi_decode(pat, 4, step=K.DEFAULT_RES // 2)
    np.testing.assert_array_equal(seq[:, :, 0], [[1, 0, 0, 0], [1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 0]])


def test_encode_decode_round_trip():                       # test.py:79-108
    seq = midi_decode(midi_encode(roll(ROLL_PLAY), step=1), 4, step=1)
    np.testing.assert_array_equal(ROLL_PLAY, seq[:, :, 0])


def test_replay_decode():                                  # test.py:110-131
    pat = one_track([midi.NoteOnEvent(tick=0, velocity=127, pitch=1), midi.NoteOnEvent(tick=0, velocity=127, pitch=3),
                     midi.NoteOffEvent(tick=1, velocity=127, pitch=1), midi.NoteOnEvent(tick=2, velocity=127, pitch=1),
                     midi.NoteOnEvent(tick=2, velocity=127, pitch=3), midi.EndOfTrackEvent(tick=1)])
    seq = midi_decode(pat, 4, step=3)
    np.testing.assert_array_equal(seq[:, :, 1], [[0, 0, 0, 0], [0, 0, 0, 1], [0, 0, 0, 0]])


def test_volume_decode():                                  # test.py:134-155
    pat = one_track([midi.NoteOnEvent(tick=0, velocity=24, pitch=0), midi.NoteOnEvent(tick=96, velocity=89, pitch=1),
                     midi.NoteOffEvent(tick=0, pitch=0), midi.NoteOffEvent(tick=48, pitch=1), midi.EndOfTrackEvent(tick=1)])
    seq = midi_decode(pat, 4, step=K.DEFAULT_RES // 2)
    np.testing.assert_array_almost_equal(seq[:, :, 2], [[24 / 127, 0, 0, 0], [24 / 127, 0, 0, 0], [0, 89 / 127, 0, 0], [0, 0, 0, 0]],
                                         decimal=5)


def test_replay_encode_decode():                           # test.py:158-193 (play channel only, as upstream)
    play = [[0, 1, 0, 1], [0, 0, 0, 1], [0, 0, 0, 1], [0, 1, 0, 1], [0, 1, 0, 1], [0, 1, 0, 1], [0, 0, 0, 0]]
    replay = [[0, 0, 0, 0]] * 4 + [[0, 0, 0, 1], [0, 1, 0, 1], [0, 0, 0, 0]]
    seq = midi_decode(midi_encode(roll(play, replay), step=2), 4, step=2)
    np.testing.assert_array_equal(play, seq[:, :, 0])


def test_replay_flag_on_unchanged_frame_is_dropped():      # midi_util.py:35,56: only changed frames are visited
    play = [[1, 0], [1, 0], [1, 1], [0, 0]]
    replay = [[0, 0], [1, 0], [1, 0], [0, 0]]
    tr = midi_encode(roll(play, replay), step=1)[0]
    kinds = [(type(e).__name__, e.tick, e.data[0] if e.data else None) for e in tr]
    # frame 1 (replay, same play vector) emits nothing; frame 2 re-articulates pitch 0 and starts pitch 1
    assert kinds == [("NoteOnEvent", 0, 0), ("NoteOffEvent", 2, 0), ("NoteOnEvent", 0, 0), ("NoteOnEvent", 0, 1),
                     ("NoteOffEvent", 1, 0), ("NoteOffEvent", 0, 1), ("EndOfTrackEvent", 0, None)]


def test_smf_write_read_round_trip(tmp_path):
    pat = midi_encode(roll(ROLL_PLAY), step=1)
    pat[0].insert(0, midi.MetaEvent(tick=0, metacommand=0x51, data=[0x07, 0xA1, 0x20]))   # tempo
    path = os.path.join(tmp_path, "t.mid")
    midi.write_midifile(path, pat)
    raw = open(path, "rb").read()
    assert raw[:4] == b"MThd" and raw[14:18] == b"MTrk"
    back = midi.read_midifile(path)
    assert back.resolution == pat.resolution and len(back) == 1
    assert [(type(a), a.tick, a.data) for a in back[0]] == [(type(a), a.tick, a.data) for a in pat[0]]
    np.testing.assert_array_equal(midi_decode(back, 4, step=1), midi_decode(pat, 4, step=1))


def test_smf_running_status_and_varlen():
    # header + one track: delta 0x81 0x00 (=128) note-on, then running status (no status byte) note-on vel 0
    body = bytes([0x81, 0x00, 0x90, 60, 100, 0x10, 60, 0, 0x00, 0xFF, 0x2F, 0x00])
    raw = b"MThd" + (6).to_bytes(4, "big") + (1).to_bytes(2, "big") * 2 + (96).to_bytes(2, "big") + b"MTrk" + len(body).to_bytes(4, "big") + body
    pat = midi.read_midifile(io.BytesIO(raw))
    ev = pat[0]
    assert (ev[0].tick, ev[0].pitch, ev[0].velocity) == (128, 60, 100)
    assert isinstance(ev[1], midi.NoteOnEvent) and (ev[1].tick, ev[1].velocity) == (16, 0)
    seq = midi_decode(pat, 128, step=16)                   # a velocity-0 note-on silences the note
    assert seq[8, 60, 0] == 1 and seq[-1, 60, 0] == 0


@pytest.mark.skipif(not os.path.isdir("/root/reference/archives/v1/long_samples"), reason="reference archives not mounted")
def test_decode_reference_archive_files():
    """The reference's own generated .mid files parse and decode to plausible piano rolls
    (resolution 4, one track, pitches inside the 36..83 range the model emits: SURVEY 2.1)."""
    d = "/root/reference/archives/v1/long_samples"
    files = sorted(f for f in os.listdir(d) if f.endswith(".mid"))
    assert files
    for f in files[:3]:
        pat = midi.read_midifile(os.path.join(d, f))
        seq = midi_decode(pat)
        assert seq.shape[1:] == (128, 3) and len(seq) >= 512
        played = np.nonzero(seq[:, :, 0].sum(0))[0]
        assert played.min() >= K.MIN_NOTE and played.max() < K.MAX_NOTE
        assert 0.01 < seq[:, K.MIN_NOTE:K.MAX_NOTE, 0].mean() < 0.15
        # re-encode / decode keeps the play channel
        again = midi_decode(midi_encode(seq, resolution=pat.resolution, step=1), step=1)
        np.testing.assert_array_equal(again[:len(seq), :, 0], seq[:, :, 0])


def test_stagger_and_load_all_on_a_tiny_corpus(tmp_path, monkeypatch):
    """dataset.stagger / load_all (dataset.py:28-76): windows every 16 steps over the left-padded roll,
    labels shifted by one step, beat one-hot of the absolute step, one style id per composer directory."""
    import dataset
    x, y = dataset.stagger([np.full((2,), i) for i in range(1, 41)], 32)
    assert len(x) == 3 and np.all(np.array(x[0]) == 0)                    # first window is pure left padding
    assert np.array(x[1])[-1, 0] == 16 and np.array(y[1])[-1, 0] == 17    # labels are one step ahead
    # a two-composer corpus written with our own encoder
    monkeypatch.chdir(tmp_path)
    rs = np.random.RandomState(0)
    groups = [["data/a/x"], ["data/b/y"]]
    for g in groups:
        os.makedirs(g[0])
        r = np.zeros((160, 128, 3))
        play = rs.rand(160, 48) < 0.06
        r[:, 36:84, 0] = play
        r[:, 36:84, 2] = play * 0.5
        midi.write_midifile(os.path.join(g[0], "piece.mid"), midi_encode(r, resolution=4, step=1))
    monkeypatch.setattr(dataset, "NUM_STYLES", 2)
    (notes, target, beat, style), (t2,) = dataset.load_all(groups, 16, 128)
    assert notes.shape[1:] == (128, 48, 3) and beat.shape[1:] == (128, 16) and style.shape[1:] == (128, 2)
    assert len(notes) == len(target) == len(beat) == len(style) and t2 is target
    assert np.array_equal(notes[:, 1:], target[:, :-1])
    assert set(np.unique(style.argmax(-1))) == {0, 1}
    k = np.nonzero(beat[2].sum(-1))[0]                                   # padded frames carry no beat
    assert np.array_equal(beat[2, k].argmax(-1), (k - k[0] + beat[2, k[0]].argmax()) % 16)
