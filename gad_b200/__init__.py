"""gad_b200 — a B200-native (sm_100a) array backend for gad's tape-based autodiff.

Only what the hot path needs lives here: `csrc/` (hand-written CUDA kernels + the C ABI declared in
include/gadcu.h) and `algebra.py`, the host-side mirror of gad's operator interface that calls the
C ABI through ctypes. There is no CPU fallback and nothing here imports `oracle/`.
"""
from .algebra import (F32, F64, BatchedGraph1, Capture, Check, Comm, Context, CuArray, CudaError, DimensionsError, EmptyError,
                      Eval, GadError, GradientId, Graph1, GraphN, InvalidError, LengthsError, MatProp, MissingGradientError,
                      MissingIdError, MissingNodeError, ReducedDimensionsError, Store, TypeMismatchError, UnsupportedError,
                      Value)

__all__ = [n for n in dir() if not n.startswith("_")]
