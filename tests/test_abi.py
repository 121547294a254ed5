"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and
exports every symbol include/deepj_b200.h declares; the host-side mirror of
the reference interface keeps the reference's names."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import music_generator_b200 as pkg
    from music_generator_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        pkg.build()
    return _lib.load()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "deepj_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dj_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from music_generator_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/deepj_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == syms


def test_version_and_error_convention(lib):
    from music_generator_b200 import _lib
    assert lib.dj_version() >= 100
    d = _lib.Dropout()
    assert lib.dj_make_dropout(7, 4, ctypes.c_float(0.5), ctypes.byref(d)) == 0
    assert d.mode == 1 and d.thr == 0x80000000 and d.scale == 2.0
    assert lib.dj_make_dropout(7, 1, ctypes.c_float(0.2), ctypes.byref(d)) == 0 and d.mode == 2
    assert lib.dj_make_dropout(7, 1, ctypes.c_float(1.5), ctypes.byref(d)) < 0       # invalid argument -> <0
    assert b"rate" in lib.dj_last_error()
    # unsupported shape is an error, never a fallback
    rc = lib.dj_lstm_scan_fwd(None, None, None, None, None, 1, 1, 512, 1, 0, 0, 0, 1, None)
    assert rc < 0


def test_peer_exchange_rejects_bad_arguments_before_any_launch(lib):
    """dj_nadam_allreduce_peer validates on the host: rank / world range, vector width, epoch numbering."""
    C = ctypes
    assert lib.dj_peer_flag_words() >= 2 * 16 + 2
    one = (C.c_void_p * 1)(C.c_void_p(16))
    sc = [C.c_float(v) for v in (1.0, 0.002, 0.9, 0.999, 1e-8, 0.9, 0.9, 0.5, 0.25, 0.001)]

    def call(rank, world, n, epoch, arr=one):
        return lib.dj_nadam_allreduce_peer(arr, arr, arr, rank, world, C.c_void_p(16), C.c_void_p(16), n, epoch,
                                           *sc, None)
    assert call(0, 17, 64, 1) < 0 and b"at most 16" in lib.dj_last_error()
    assert call(2, 2, 64, 1) < 0
    assert call(0, 1, 66, 1) < 0 and b"multiple of 4" in lib.dj_last_error()
    assert call(0, 1, 64, 0) < 0 and b"epoch" in lib.dj_last_error()
    assert lib.dj_nadam_allreduce_peer(None, one, one, 0, 1, C.c_void_p(16), C.c_void_p(16), 64, 1, *sc, None) < 0
    assert lib.dj_peer_open(None, None) < 0 and lib.dj_peer_close(None) < 0 and lib.dj_peer_free(None) < 0


def test_round2_entry_points_validate_arguments_on_the_host(lib):
    """The entry points added in round 2 reject bad arguments before any CUDA call (so this runs without a GPU)."""
    import ctypes as C
    one = C.c_void_p(16)
    # fused two-layer generation scan: NULL pointers, ragged sequence counts, more clusters than can be co-resident
    P13 = [one] * 13
    assert lib.dj_lstm_scan_tc_gen2(None, *P13, 1.0 / 1024, one, 48, 128, 1, None) < 0
    assert b"NULL" in lib.dj_last_error()
    assert lib.dj_lstm_scan_tc_gen2(one, *P13, 1.0 / 1024, one, 50, 128, 1, None) < 0
    assert lib.dj_lstm_scan_tc_gen2(one, *P13, 1.0 / 1024, one, 48 * 3, 128, 1, None) < 0
    assert b"at most 2 sequences" in lib.dj_last_error()
    assert lib.dj_lstm_scan_tc_gen2(one, *P13, 0.0, one, 48, 128, 1, None) < 0
    # deterministic-reduction workspace: registering, replacing and unregistering is host-side bookkeeping
    st = C.c_void_p(0x1000)
    assert lib.dj_set_reduce_workspace(st, None, 64) < 0            # a size without a buffer
    assert lib.dj_set_reduce_workspace(st, one, -1) < 0
    assert lib.dj_set_reduce_workspace(st, one, 1 << 20) == 0
    assert lib.dj_set_reduce_workspace(st, one, 1 << 21) == 0       # same stream: replaced, not a second slot
    assert lib.dj_set_reduce_workspace(st, None, 0) == 0
    for i in range(8):                                              # eight streams fit, the ninth is refused
        assert lib.dj_set_reduce_workspace(C.c_void_p(0x2000 + i), one, 16) == 0
    assert lib.dj_set_reduce_workspace(C.c_void_p(0x3000), one, 16) < 0
    for i in range(8):
        assert lib.dj_set_reduce_workspace(C.c_void_p(0x2000 + i), None, 0) == 0


def test_dropout_key_matches_numpy_twin(lib):
    from music_generator_b200 import _lib
    import helpers
    for seed, site in ((7, 4), (123456789012345, 11), (0, 1)):
        d = _lib.make_dropout(seed, site, 0.5)
        assert d.key == int(helpers.site_key(seed, site))


def test_reference_entry_points_exist():
    import constants
    assert constants.NUM_STYLES == 23 and constants.NUM_NOTES == 48 and constants.SEQ_LEN == 128
    assert constants.TIME_AXIS_UNITS == 256 and constants.NOTE_AXIS_UNITS == 128
    import dataset, util
    assert util.one_hot(2, 4).tolist() == [0, 0, 1, 0]
    assert dataset.compute_beat(17, 16)[1] == 1
    g = dataset.compute_genre(2)
    assert np.count_nonzero(g) == 14 and abs(g.sum() - 1) < 1e-12
    assert dataset.unclamp_midi(np.ones((5, 48, 3))).shape == (5, 84, 3)
    src = open(os.path.join(ROOT, "model.py")).read()
    assert "def build_models(time_steps=SEQ_LEN, input_dropout=0.2, dropout=0.5" in src


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from music_generator_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine()
