"""options_model_b200 -- B200-native American-option Monte Carlo engine (path simulation + LSM).

Layout:
  csrc/        hand-written sm_100a CUDA kernels + the extern "C" boundary (include/optmc.h)
  _lib.py      ctypes binding of liboptmc.so (fails loudly when the library or a GPU is missing)
  engine.py    thin Python handle over the C ABI (device buffers are torch tensors)
  compat.py    mirror of the reference's Python call signatures (AdvancedOptionPricer, ...)
  sharded.py   path-sharded multi-GPU sweep (per-date Gram all-reduce over torch.distributed)
"""
from . import _lib  # noqa: F401
from ._lib import OptmcError, library_path, load_library  # noqa: F401

__all__ = ["OptmcError", "library_path", "load_library"]
