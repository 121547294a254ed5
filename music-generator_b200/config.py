"""Model hyper-parameters of the DeepJ hot path (reference constants.py:42-77)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple


@dataclass(frozen=True)
class ModelConfig:
    num_styles: int = 23
    num_notes: int = 48
    note_units: int = 3
    notes_per_bar: int = 16
    seq_len: int = 128
    octave_units: int = 64
    style_units: int = 64
    time_axis_units: int = 256
    note_axis_units: int = 128

    @property
    def feat0(self) -> int:
        return 1 + 12 + 1 + self.octave_units + self.notes_per_bar

    def layers(self) -> List[dict]:
        """The four LSTM layers in execution order with their dropout site ids
        (D1..D12 numbering of SURVEY.md 8a)."""
        ut, un = self.time_axis_units, self.note_axis_units
        return [
            dict(name="time0", axis="time", F=self.feat0, U=ut, site_in=None, site_sp=5, site_out=6),
            dict(name="time1", axis="time", F=ut, U=ut, site_in=6, site_sp=7, site_out=8),
            dict(name="note0", axis="note", F=ut + self.note_units, U=un, site_in=8, site_sp=9, site_out=10),
            dict(name="note1", axis="note", F=un, U=un, site_in=10, site_sp=11, site_out=12),
        ]


def param_shapes(cfg: ModelConfig) -> Dict[str, Tuple[int, ...]]:
    """28 tensors, Keras layouts, model.py creation order (Dense [in,out],
    Conv1D [k,in,out], LSTM kernel [in,4U] / recurrent [U,4U] / bias [4U], gate
    blocks i,f,c,o)."""
    s: Dict[str, Tuple[int, ...]] = {}
    s["style.W"] = (cfg.num_styles, cfg.style_units)
    s["style.b"] = (cfg.style_units,)
    s["conv.W"] = (24, cfg.note_units, cfg.octave_units)
    s["conv.b"] = (cfg.octave_units,)
    for L in cfg.layers():
        n, f, u = L["name"], L["F"], L["U"]
        s[f"{n}.sd.W"] = (cfg.style_units, f)
        s[f"{n}.sd.b"] = (f,)
        s[f"{n}.lstm.W"] = (f, 4 * u)
        s[f"{n}.lstm.U"] = (u, 4 * u)
        s[f"{n}.lstm.b"] = (4 * u,)
    s["note_dense.W"] = (cfg.note_axis_units, 2)
    s["note_dense.b"] = (2,)
    s["volume_dense.W"] = (cfg.note_axis_units, 1)
    s["volume_dense.b"] = (1,)
    return s


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m
