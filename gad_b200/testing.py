"""gad's `testing` helpers over device arrays — SURVEY §8(f) row 4 (gad/src/arrayfire.rs:114-164).

`estimate_gradient` is the reference's central-difference estimate of d dot(f(x), direction) / dx, one element of x at
a time; `assert_almost_all_equal` its L-infinity comparison. With these every array test of the reference
(gad/tests/*.rs `af` modules) replays verbatim against this backend (tests/test_gpu_reference_tests.py).
"""
from __future__ import annotations

from typing import Callable

import numpy as np

from .algebra import Context, CuArray, Eval, Value


def _host(v) -> np.ndarray:
    d = v.data() if isinstance(v, Value) else v
    return d.numpy() if isinstance(d, CuArray) else np.asarray(d)


def estimate_gradient(ctx: Context, x: np.ndarray, direction: np.ndarray, epsilon: float, f: Callable) -> np.ndarray:
    """arrayfire.rs:120-150. `f(eval_algebra, value) -> value` is evaluated on the device for x +/- epsilon in every
    coordinate. The dot with `direction` is taken in f64 on the host so the difference quotient is limited by the
    f32 rounding of f itself (the reference's CPU runs behave the same way)."""
    ev = Eval(ctx)
    v = np.asarray(x, dtype=np.float32).copy()
    grad = np.zeros_like(v)
    flat, gflat = v.reshape(-1, order="F"), grad.reshape(-1, order="F")
    dh = np.asarray(direction, dtype=np.float64).reshape(-1, order="F")

    def y(values):
        out = _host(f(ev, ev.constant(values.reshape(v.shape, order="F"))))
        return float(np.dot(out.reshape(-1, order="F").astype(np.float64), dh))

    for i in range(flat.size):
        x0 = flat[i]
        flat[i] = x0 + epsilon
        y2 = y(flat)
        flat[i] = x0 - epsilon
        y1 = y(flat)
        gflat[i] = (y2 - y1) / (epsilon + epsilon)
        flat[i] = x0
    return gflat.reshape(v.shape, order="F")


def assert_almost_all_equal(v1, v2, precision: float) -> None:
    """arrayfire.rs:153-163: same dims and max |v1 - v2| < precision."""
    a, b = np.squeeze(_host(v1)), np.squeeze(_host(v2))
    assert a.shape == b.shape, (a.shape, b.shape)
    d = float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) if a.size else 0.0
    assert d < precision, f"L-infinity distance {d} >= {precision}"
