"""ctypes binding of libdeepj_sm100.so (include/deepj_b200.h).

The library is the product path: there is no CPU fallback.  `load()` raises if
the shared object is missing; every wrapper raises RuntimeError with
dj_last_error() when an entry point returns non-zero.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdeepj_sm100.so")
BUILD_SCRIPT = os.path.join(_HERE, "csrc", "build.sh")

DJ_F32, DJ_BF16, DJ_F16 = 0, 1, 2


class Dropout(C.Structure):
    """struct dj_dropout"""
    _fields_ = [("key", C.c_uint32), ("thr", C.c_uint32), ("scale", C.c_float), ("mode", C.c_int32),
                ("key_ptr", C.c_void_p)]


_p, _i, _i64, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double

# name -> (restype, argtypes); must list every symbol include/deepj_b200.h declares
SIGNATURES = {
    "dj_version": (_i, []),
    "dj_last_error": (C.c_char_p, []),
    "dj_set_reduce_workspace": (_i, [_p, _p, _i64]),
    "dj_make_dropout": (_i, [C.c_uint64, _i, _f, C.POINTER(Dropout)]),
    "dj_dropout_site_key": (C.c_uint32, [C.c_uint64, _i]),
    "dj_dropout_mask_materialize": (_i, [Dropout, _i64, _i, _p, _p]),
    "dj_style_fwd": (_i, [_p, _i64, _i64, _i, _i, _i, _p, _p, _i, C.POINTER(_p), C.POINTER(_p),
                          C.POINTER(_i), _p, C.POINTER(_p), _p]),
    "dj_frontend_fwd": (_i, [_p, _i64, _p, _i64, _i, _i, _p, _p, _p, Dropout, Dropout, Dropout, Dropout,
                             _p, _p, _i, _i, _p]),
    "dj_layer_input": (_i, [_p, _i, _i64, _i64, Dropout, _p, _i, Dropout, _p, _i64, Dropout, _i, _i, _p, _p,
                            _i, _i, _p]),
    "dj_gemm_simt": (_i, [_p, _i, _i64, _i64, _p, _i, _i64, _i64, _p, _i64, _p, _i, _i, _i, _i, _i64,
                          _i64, _p]),
    "dj_gate_gemm_bf16": (_i, [_p, _i64, _p, _i64, _p, _i64, _p, _i, _i, _i, _p]),
    "dj_gate_gemm_16": (_i, [_p, _p, _i, _i64, _p, _p, _i, _i64, _p, _i64, _p, _i, _i, _i, _p]),
    "dj_gate_gemm_16s": (_i, [_p, _p, _i, _i64, _p, _p, _i, _i64, _p, _i64, _p, _f, _i, _i, _i, _p]),
    "dj_wgrad_gemm_bf16": (_i, [_p, _i64, _p, _i64, _p, _i64, _i, _i, _i64, _p]),
    "dj_wgrad_gemm_16": (_i, [_p, _i, _i64, _p, _i, _i64, _p, _i64, _i, _i, _i64, _p]),
    "dj_cast_bf16": (_i, [_p, _i, _i, _p, _i, _i, _p]),
    "dj_half_to_bf16_inplace": (_i, [_p, _i64, _p]),
    "dj_cast16_multi": (_i, [_i, C.POINTER(_p), C.POINTER(_i), C.POINTER(_i), C.POINTER(_p), C.POINTER(_p),
                             C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_f), _p]),
    "dj_lstm_scan_tc_infer": (_i, [_p, _p, _p, _p, _p, _p, _f, _i, _i, _i, _i, _i64, _i64, _i64, _i, _p]),
    "dj_lstm_scan_tc_gen2": (_i, [_p] * 14 + [_f, _p, _i, _i, _i, _p]),
    "dj_lstm_scan_fwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i64, _i64, _i64, _i, _p]),
    "dj_lstm_scan_tc_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i64, _i64, _i64, _i, _p]),
    "dj_lstm_scan_tc_bwd": (_i, [_p, _p, _p, _i64, Dropout, _p, _p, _p, _i, _i, _i, _i, _i64, _i64, _i64, _i, _p]),
    "dj_lstm_scan_bwd": (_i, [_p, _p, _p, _i64, Dropout, _p, _p, _i, _p, _i, _i, _i, _i, _i64, _i64,
                              _i64, _i, _p]),
    "dj_head_partials_size": (_i64, [_i]),
    "dj_head_loss": (_i, [_p, _i, Dropout, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "dj_head_finalize": (_i, [_p, _i, _p, _p, _p, _p, _p, _p]),
    "dj_style_bwd_reduce": (_i, [_p, _i64, _i, _p, Dropout, _i, _p, _p]),
    "dj_colsum": (_i, [_p, _i64, _i64, _i, _p, _i, _p]),
    "dj_conv_bwd": (_i, [_p, _i64, _i, _i, _p, _p, Dropout, Dropout, _p, _i64, _p, _p, _p]),
    "dj_nadam_step": (_i, [_p, _p, _p, _p, _i64, _f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _p]),
    "dj_nadam_step_dev": (_i, [_p, _p, _p, _p, _i64, _p, _p]),
    "dj_nadam_allreduce_peer_dev": (_i, [C.POINTER(_p), C.POINTER(_p), C.POINTER(_p), _i, _i, _p, _p, _i64, _p, _p, _p]),
    "dj_peer_flag_words": (_i64, []),
    "dj_peer_alloc": (_i, [_i64, C.POINTER(_p), C.c_char_p]),
    "dj_peer_open": (_i, [C.c_char_p, C.POINTER(_p)]),
    "dj_peer_close": (_i, [_p]),
    "dj_peer_free": (_i, [_p]),
    "dj_nadam_allreduce_peer": (_i, [C.POINTER(_p), C.POINTER(_p), C.POINTER(_p), _i, _i, _p, _p, _i64, C.c_uint32,
                                     _f, _f, _f, _f, _f, _f, _f, _f, _f, _f, _p]),
    "dj_gen_sample": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _p, _p, _i, _p, _p, _d,
                           _i, _p, _p, _p, _p]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a (works without a GPU)."""
    r = subprocess.run(["bash", BUILD_SCRIPT], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("nvcc build of libdeepj_sm100.so failed")
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with music-generator_b200/csrc/build.sh "
            "(__graft_entry__.build()).  There is no CPU fallback for the DeepJ hot path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().dj_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what or 'deepj'} failed (rc={rc}): {msg}")


def make_dropout(seed: int, site: int, rate: float) -> Dropout:
    d = Dropout()
    check(load().dj_make_dropout(C.c_uint64(seed & (2 ** 64 - 1)), site, float(rate), C.byref(d)),
          "dj_make_dropout")
    return d


NO_DROPOUT = Dropout(0, 0, 1.0, 0, None)
