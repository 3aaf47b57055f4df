"""GPU: the f64 tensor-core GEMM (mma.sync m8n8k4 DMMA, csrc/gemm_dmma.cu) against numpy: all four major-ness
combinations, ragged sizes (edge tiles), batch broadcast, and the accumulate epilogue through the tape.
Tolerance: every DMMA rounds a 4-term dot product into the f64 accumulator, K/4 roundings of <= 2^-53 of the running
sum each plus the products' own roundings: |err| <= (K/4 + 4) * 2^-52 * sum_k |a_k||b_k|."""
import numpy as np
import pytest

import gad_b200 as gb
from helpers import to_np

pytestmark = pytest.mark.gpu


def _check(ctx, M, N, K, t1, t2, seed=0):
    rng = np.random.default_rng(seed)
    sa, sb = ((K, M) if t1 else (M, K)), ((N, K) if t2 else (K, N))
    a, b = rng.standard_normal(sa), rng.standard_normal(sb)
    ev = gb.Eval(ctx)
    got = to_np(ev.matmul(ev.constant(a), ev.constant(b), gb.MatProp(t1), gb.MatProp(t2)))[:, :, 0, 0]
    assert got.dtype == np.float64
    opA, opB = (a.T if t1 else a), (b.T if t2 else b)
    want = opA @ opB
    S = np.abs(opA) @ np.abs(opB)
    err = np.abs(got - want)
    bound = (K / 4 + 4) * 2.0 ** -52 * S
    assert np.all(err <= bound), (M, N, K, t1, t2, float((err / bound).max()))


@pytest.mark.parametrize("t1,t2", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(64, 64, 4), (128, 128, 16), (256, 384, 64), (200, 136, 52), (129, 65, 7), (1024, 512, 333)])
def test_dmma_matches_numpy(ctx, M, N, K, t1, t2):
    _check(ctx, M, N, K, t1, t2)


def test_dmma_batch_broadcast(ctx):
    """matmul over dims 2,3 with a broadcast right operand (matrix.rs:108-114)."""
    rng = np.random.default_rng(1)
    a = rng.standard_normal((96, 40, 3, 2))
    b = rng.standard_normal((40, 72))
    ev = gb.Eval(ctx)
    got = to_np(ev.matmul_nn(ev.constant(a), ev.constant(b)))
    assert got.shape == (96, 72, 3, 2)
    for i in range(3):
        for j in range(2):
            assert np.allclose(got[:, :, i, j], a[:, :, i, j] @ b, rtol=1e-13, atol=1e-12)


def test_dmma_through_the_tape_with_accumulation(ctx):
    """f64 matmul VJPs (dA = G.B^T, dB = A^T.G) and `current + new` in the GEMM epilogue (w used twice)."""
    rng = np.random.default_rng(2)
    M, K, N = 192, 160, 128
    x, w = rng.standard_normal((M, K)), rng.standard_normal((K, N))
    g = gb.Graph1(ctx)
    vx, vw = g.variable(x), g.variable(w)
    y = g.add(g.matmul_nn(vx, vw), g.matmul_nn(vx, vw))
    seed = rng.standard_normal((M, N))
    store = g.evaluate_gradients_once(y.gid(), seed)
    gw = to_np(store.get(vw.gid()))[:, :, 0, 0]
    gx = to_np(store.get(vx.gid()))[:, :, 0, 0]
    assert np.allclose(gw, 2.0 * (x.T @ seed), rtol=1e-12, atol=1e-11)
    assert np.allclose(gx, 2.0 * (seed @ w.T), rtol=1e-12, atol=1e-11)
