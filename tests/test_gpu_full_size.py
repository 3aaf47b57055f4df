"""GPU: the BASELINE.json configurations at their FULL sizes, checked through properties that do not need a full-size
CPU answer (the oracle only sees samples it finishes in seconds):
  C2  2^20 lanes x Rosenbrock-64: sampled lanes are bit-identical to the oracle's scalar tape; closed form at x = 1.
  C3  64 Mi elements: sampled gradients against the f64 closed form; the sweep is linear in its seed (a power of two
      scales every gradient exactly); z against an f64 sum of a strided sample's closed form is left to the small tests.
  C4  B = 8192, D = 4096: the sweep is linear in its seed (exactly, for a power of two); the gradients of the two halves
      of the batch add up to the full-batch gradients (the data-parallel seam) within the accumulation tolerance; the
      loss is the sum of the halves' losses.
"""
import numpy as np
import pytest

import gad_b200 as gb
from gad_b200 import workloads as wl
from helpers import to_np

pytestmark = pytest.mark.gpu


def test_c2_full_size_sampled_lanes_match_the_scalar_tape(ctx, oracle):
    lanes, N = 1 << 20, 64
    cols = [ctx.random((lanes,), 200 + i, dtype=gb.F64, a=-1.2, b=1.2) for i in range(N)]
    g, f, grads, store = wl.rosenbrock_batched_gradients(ctx, cols)
    idx = np.random.default_rng(0).choice(lanes, 2048, replace=False)
    x = np.stack([to_np(c)[:, 0, 0, 0][idx] for c in cols])
    got = np.stack([to_np(c)[:, 0, 0, 0][idx] for c in grads])
    _, ograd, _ = oracle.rosenbrock_batch(x, nthreads=8)
    assert np.array_equal(got, ograd)
    # re-run at the stationary point x = (1, ..., 1): every gradient is exactly zero
    ones = [ctx.array(np.ones(lanes)) for _ in range(N)]
    at_one = store.rerun(ones)
    assert all(not np.any(to_np(c)) for c in at_one)


def test_c3_full_size_closed_form_and_seed_linearity(ctx):
    n = 64 * 1024 * 1024
    x = ctx.random((n,), 3, a=-1.0, b=1.0)
    y = ctx.random((n,), 4, a=0.5, b=2.0)
    g = gb.Graph1(ctx)
    vx, vy, z = wl.c3_forward(g, x, y)
    s1 = g.evaluate_gradients(z.gid(), 1.0)
    s4 = g.evaluate_gradients(z.gid(), 4.0)
    gx1, gy1 = s1.get(vx.gid()).numpy().reshape(-1), s1.get(vy.gid()).numpy().reshape(-1)
    gx4, gy4 = s4.get(vx.gid()).numpy().reshape(-1), s4.get(vy.gid()).numpy().reshape(-1)
    # dz/dy = seed * exp(x) / y scales exactly with a power-of-two seed; dz/dx = seed + (seed * log y) * exp(x) too
    assert np.array_equal(gy4, 4.0 * gy1) and np.array_equal(gx4, 4.0 * gx1)
    idx = np.random.default_rng(1).choice(n, 1 << 16, replace=False)
    hx, hy = x.numpy().reshape(-1)[idx].astype(np.float64), y.numpy().reshape(-1)[idx].astype(np.float64)
    eps = np.finfo(np.float32).eps
    want_gy = np.exp(hx) / hy
    assert np.all(np.abs(gy1[idx] - want_gy) <= 8 * eps * np.abs(want_gy))
    prod = np.abs(np.exp(hx) * np.log(hy))
    assert np.all(np.abs(gx1[idx] - (np.exp(hx) * np.log(hy) + 1.0)) <= 6 * eps * (prod + 1.0))


def test_c4_full_size_seed_linearity_and_batch_additivity(ctx):
    B, D, L = 8192, 4096, 3
    X = ctx.random((B, D), 5, normal=True, a=0.0, b=1.0)
    T = ctx.random((B, D), 9, normal=True, a=0.0, b=1.0)
    Ws = [ctx.random((D, D), 6 + l, normal=True, a=0.0, b=1.0 / np.sqrt(D)) for l in range(L)]
    bs = [ctx.array(np.full((1, D), 0.01, np.float32)) for _ in range(L)]
    # seed linearity: one tape, two sweeps
    g = gb.Graph1(ctx)
    loss, vW, vb = wl.mlp_forward(g, X, Ws, bs, T)
    s1 = g.evaluate_gradients(loss.gid(), 1.0)
    s2 = g.evaluate_gradients(loss.gid(), 2.0)
    for v in vW + vb:
        a, b = s1.get(v.gid()).numpy(), s2.get(v.gid()).numpy()
        assert np.array_equal(b, 2.0 * a)
    full_W = [s1.get(v.gid()).numpy()[:, :, 0, 0].astype(np.float64) for v in vW]
    full_loss = loss.data().item()
    del s2, g
    # the data-parallel seam at full size: rows [0, B/2) and [B/2, B)
    hX, hT = X.numpy()[:, :, 0, 0], T.numpy()[:, :, 0, 0]
    parts = []
    for sl in (slice(0, B // 2), slice(B // 2, B)):
        l_, gW, gb_ = wl.mlp_step(ctx, ctx.array(np.ascontiguousarray(hX[sl])), Ws, bs, ctx.array(np.ascontiguousarray(hT[sl])))
        parts.append((l_.item(), [w.numpy()[:, :, 0, 0].astype(np.float64) for w in gW]))
    assert abs(parts[0][0] + parts[1][0] - full_loss) <= 1e-5 * abs(full_loss)
    for l in range(L):
        summed = parts[0][1][l] + parts[1][1][l]
        scale = np.max(np.abs(full_W[l]))
        # f32 accumulation over 8192 rows in two different associations
        assert np.max(np.abs(summed - full_W[l])) <= 2e-5 * scale * np.sqrt(B) / 8, l
