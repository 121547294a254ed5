"""deepj-b200: B200-native (sm_100a) implementation of the DeepJ biaxial-LSTM hot path.

Importable as `music_generator_b200` through the shim at the repository root
(the directory name carries the reference's hyphen).
"""
from .config import ModelConfig, param_shapes  # noqa: F401
from . import _lib  # noqa: F401


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu into libdeepj_sm100.so (sm_100a, needs nvcc only)."""
    return _lib.build(verbose)
