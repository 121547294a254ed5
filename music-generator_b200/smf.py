"""Minimal Standard MIDI File object model, reader and writer.

The reference depends on the third-party `python-midi` package (`import midi`,
reference midi_util.py:4, README.md:10-13), which is not installable here.  This
module provides the small slice of that API the reference's call sites use --
`Pattern(resolution=)`, `Track`, `NoteOnEvent(tick=, velocity=, pitch=)`,
`NoteOffEvent`, `EndOfTrackEvent`, `.tick` (delta ticks), `.data`,
`read_midifile`, `write_midifile` -- implemented from the SMF 1.0 specification.
"""
from __future__ import annotations

import struct
from typing import BinaryIO, List


class Event:
    status = None          # channel-voice status nibble (0x80, 0x90, ...) or None for meta events
    name = "Event"

    def __init__(self, tick: int = 0, data=None, channel: int = 0):
        self.tick = int(tick)
        self.data = list(data) if data is not None else []
        self.channel = int(channel)

    def __repr__(self):
        return f"{self.name}(tick={self.tick}, data={self.data})"


class _NoteEvent(Event):
    def __init__(self, tick: int = 0, pitch: int = 0, velocity: int = 0, channel: int = 0, data=None):
        if data is not None:
            pitch, velocity = data
        super().__init__(tick, [int(pitch), int(velocity)], channel)

    pitch = property(lambda s: s.data[0], lambda s, v: s.data.__setitem__(0, int(v)))
    velocity = property(lambda s: s.data[1], lambda s, v: s.data.__setitem__(1, int(v)))


class NoteOnEvent(_NoteEvent):
    status, name = 0x90, "NoteOnEvent"


class NoteOffEvent(_NoteEvent):
    status, name = 0x80, "NoteOffEvent"


class ChannelEvent(Event):
    """Any other channel-voice message (kept so files survive a read/write round trip)."""
    name = "ChannelEvent"

    def __init__(self, tick=0, status=0xB0, data=None, channel=0):
        super().__init__(tick, data, channel)
        self.status = status


class MetaEvent(Event):
    name = "MetaEvent"

    def __init__(self, tick=0, metacommand=0, data=None):
        super().__init__(tick, data)
        self.metacommand = metacommand


class EndOfTrackEvent(MetaEvent):
    name = "EndOfTrackEvent"

    def __init__(self, tick=0, data=None):
        super().__init__(tick, 0x2F, data or [])


class SysexEvent(Event):
    name = "SysexEvent"


class Track(list):
    pass


class Pattern(list):
    def __init__(self, tracks=(), resolution: int = 220, format: int = 1):
        super().__init__(tracks)
        self.resolution = resolution
        self.format = format


# ---------------------------------------------------------------- variable-length quantities
def _write_varlen(v: int) -> bytes:
    out = [v & 0x7F]
    v >>= 7
    while v:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    return bytes(reversed(out))


def _read_varlen(buf: bytes, pos: int):
    v = 0
    while True:
        b = buf[pos]; pos += 1
        v = (v << 7) | (b & 0x7F)
        if not b & 0x80:
            return v, pos


_CHANNEL_DATA_LEN = {0x80: 2, 0x90: 2, 0xA0: 2, 0xB0: 2, 0xC0: 1, 0xD0: 1, 0xE0: 2}


def _parse_track(buf: bytes) -> Track:
    track, pos, running = Track(), 0, None
    while pos < len(buf):
        tick, pos = _read_varlen(buf, pos)
        b = buf[pos]
        if b == 0xFF:                                  # meta event
            cmd = buf[pos + 1]
            n, p2 = _read_varlen(buf, pos + 2)
            data = list(buf[p2:p2 + n]); pos = p2 + n
            ev = EndOfTrackEvent(tick, data) if cmd == 0x2F else MetaEvent(tick, cmd, data)
            track.append(ev)
            if cmd == 0x2F:
                break
            continue
        if b in (0xF0, 0xF7):                          # sysex
            n, p2 = _read_varlen(buf, pos + 1)
            track.append(SysexEvent(tick, list(buf[p2:p2 + n]))); pos = p2 + n
            continue
        if b & 0x80:
            running = b; pos += 1
        elif running is None:
            raise ValueError("SMF data byte without running status")
        st, ch = running & 0xF0, running & 0x0F
        n = _CHANNEL_DATA_LEN[st]
        data = list(buf[pos:pos + n]); pos += n
        if st == 0x90:
            track.append(NoteOnEvent(tick, data[0], data[1], ch))
        elif st == 0x80:
            track.append(NoteOffEvent(tick, data[0], data[1], ch))
        else:
            track.append(ChannelEvent(tick, st, data, ch))
    return track


def read_midifile(f) -> Pattern:
    own = isinstance(f, (str, bytes))
    fh: BinaryIO = open(f, "rb") if own else f
    try:
        raw = fh.read()
    finally:
        if own:
            fh.close()
    if raw[:4] != b"MThd":
        raise ValueError("not a Standard MIDI File")
    hlen, fmt, ntrk, division = struct.unpack(">IHHH", raw[4:14])
    if division & 0x8000:
        raise ValueError("SMPTE time division is not supported")
    pat, pos = Pattern(resolution=division, format=fmt), 8 + hlen
    for _ in range(ntrk):
        if raw[pos:pos + 4] != b"MTrk":
            raise ValueError("missing MTrk chunk")
        (n,) = struct.unpack(">I", raw[pos + 4:pos + 8])
        pat.append(_parse_track(raw[pos + 8:pos + 8 + n]))
        pos += 8 + n
    return pat


def _encode_track(track: List[Event]) -> bytes:
    out = bytearray()
    for ev in track:
        out += _write_varlen(ev.tick)
        if isinstance(ev, MetaEvent):
            out += bytes([0xFF, ev.metacommand]) + _write_varlen(len(ev.data)) + bytes(ev.data)
        elif isinstance(ev, SysexEvent):
            out += bytes([0xF0]) + _write_varlen(len(ev.data)) + bytes(ev.data)
        else:
            if any(not 0 <= d <= 0x7F for d in ev.data):
                raise ValueError(f"MIDI data byte out of range 0..127 in {type(ev).__name__}: {list(ev.data)}")
            out += bytes([ev.status | ev.channel]) + bytes(ev.data)
    if not track or not isinstance(track[-1], EndOfTrackEvent):
        out += b"\x00\xff\x2f\x00"
    return bytes(out)


def write_midifile(f, pattern: Pattern) -> None:
    own = isinstance(f, (str, bytes))
    fh: BinaryIO = open(f, "wb") if own else f
    try:
        fh.write(b"MThd" + struct.pack(">IHHH", 6, getattr(pattern, "format", 1), len(pattern), pattern.resolution))
        for tr in pattern:
            body = _encode_track(tr)
            fh.write(b"MTrk" + struct.pack(">I", len(body)) + body)
    finally:
        if own:
            fh.close()
