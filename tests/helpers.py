"""Shared helpers: run one generic gad program on the CUDA backend and on the oracle, compare."""
import numpy as np

import gad_b200 as gb


def to_np(v):
    """Host copy of a value's data from either backend (arrays as 4-D column-major, scalars as float)."""
    if isinstance(v, gb.Value):
        d = v.data()
        return d.item() if d.is_scalar else d.numpy()
    if isinstance(v, gb.CuArray):
        return v.item() if v.is_scalar else v.numpy()
    return v.data()


def _ordered(x):
    """Map float bits to integers that are monotonic in the float value (so ulp distance = difference)."""
    it = np.int32 if x.dtype == np.float32 else np.int64
    i = x.view(it).astype(np.int64)
    mask = np.int64(0x7FFFFFFF) if x.dtype == np.float32 else np.int64(0x7FFFFFFFFFFFFFFF)
    return np.where(i < 0, -(i & mask), i)


def ulp_diff(a, b):
    """Max distance in units in the last place between two same-dtype float arrays."""
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.dtype == b.dtype and a.shape == b.shape, (a.dtype, b.dtype, a.shape, b.shape)
    if a.size == 0:
        return 0
    return int(np.max(np.abs(_ordered(a) - _ordered(b))))


def rand(shape, seed, lo=0.0, hi=1.0, dtype=np.float32):
    rng = np.random.default_rng(seed)
    return np.asarray(rng.uniform(lo, hi, size=shape), dtype=dtype)
