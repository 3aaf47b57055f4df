"""GPU: the tcgen05 + TMA 3xTF32 GEMM against f64 numpy, all four major-ness combinations (the forward
product, dA = G.B^T and dB = A^T.G feed K-major / MN-major operands directly), accumulate epilogue."""
import numpy as np
import pytest

import gad_b200 as gb
from helpers import to_np

pytestmark = pytest.mark.gpu


def _check(ctx, M, N, K, t1, t2, seed=0):
    rng = np.random.default_rng(seed)
    sa, sb = ((K, M) if t1 else (M, K)), ((N, K) if t2 else (K, N))
    a = rng.standard_normal(sa).astype(np.float32)
    b = rng.standard_normal(sb).astype(np.float32)
    ev = gb.Eval(ctx)
    got = to_np(ev.matmul(ev.constant(a), ev.constant(b), gb.MatProp(t1), gb.MatProp(t2)))[:, :, 0, 0].astype(np.float64)
    A, B = a.astype(np.float64), b.astype(np.float64)
    opA, opB = (A.T if t1 else A), (B.T if t2 else B)
    want = opA @ opB
    # tolerance (DESIGN.md "GEMM numerics"): 3xTF32 drops a_lo.b_lo and rounds the two lo parts,
    # <= 3 * 2^-22 |a||b| per product; the f32 accumulator in TMEM is rounded once per tcgen05.mma, i.e.
    # 3 * K/8 times, each <= 2^-23 of the running sum (the tensor core truncates). Both are bounded by
    # multiples of S = sum_k |a_k||b_k|; a plain f32 GEMM carries K * 2^-24 S from its accumulation alone.
    S = np.abs(opA) @ np.abs(opB)
    bound = (2.0 ** -20 + 3 * (K / 8) * 2.0 ** -23) * S
    err = np.abs(got - want)
    assert np.all(err <= bound), (M, N, K, t1, t2, float(err.max()), float((err / bound).max()))
    # and the typical error stays at f32 level: rms(err / S) within 2^-20 (single-pass TF32 would be ~2^-12)
    assert float(np.sqrt(np.mean((err / S) ** 2))) <= 2.0 ** -20
    return float((err / (2.0 ** -20 * S)).max())


@pytest.mark.parametrize("t1,t2", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 256), (256, 128, 64), (128, 256, 96), (384, 512, 160), (1024, 1024, 1024)])
def test_gemm_tc_matches_f64(ctx, M, N, K, t1, t2):
    _check(ctx, M, N, K, t1, t2)


def test_gemm_tc_is_far_more_accurate_than_single_tf32(ctx):
    """3xTF32 stays within a few 2^-20 of the |a||b| sum at K=2048; a single-pass TF32 product is ~2^-11."""
    worst = _check(ctx, 512, 512, 2048, False, False, seed=3)
    assert worst < 8.0


def test_gemm_tc_accumulate_epilogue(ctx):
    """dW accumulation `current + new` fused into the GEMM epilogue: two contributions into one gradient."""
    rng = np.random.default_rng(1)
    M = N = K = 256
    x = rng.standard_normal((M, K)).astype(np.float32)
    w = rng.standard_normal((K, N)).astype(np.float32)
    g = gb.Graph1(ctx)
    vw = g.variable(w)
    cx = g.constant(x)
    y = g.add(g.matmul_nn(cx, vw), g.matmul_nn(cx, vw))  # w used twice: second dW accumulates
    seed = rng.standard_normal((M, N)).astype(np.float32)
    store = g.evaluate_gradients_once(y.gid(), seed)
    got = to_np(store.get(vw.gid()))[:, :, 0, 0].astype(np.float64)
    want = 2.0 * (x.astype(np.float64).T @ seed.astype(np.float64))
    bound = 2.0 ** -19 * (np.abs(x.astype(np.float64)).T @ np.abs(seed.astype(np.float64)))
    assert np.all(np.abs(got - want) <= bound)


# ---- shapes that are not multiples of the tile: zero-filled edge tiles, predicated epilogue -----------------------
@pytest.mark.parametrize("t1,t2", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(1000, 1000, 1000), (300, 4100, 777), (4097, 100, 333), (77, 5000, 99), (257, 513, 4099)])
def test_gemm_tc_ragged_shapes_match_f64(ctx, M, N, K, t1, t2):
    launches = ctx.launches()
    _check(ctx, M, N, K, t1, t2, seed=M + K)
    assert ctx.launches() - launches == 3  # two operand splits + ONE tensor-core GEMM (no CUDA-core fallback pieces)


def test_gemm_tc_ragged_is_exact_on_small_integers(ctx):
    """Integer-valued operands: every partial product and sum is exact in TF32 x TF32 -> f32, so any element that
    came from the wrong place (edge tile, padded column, k tail) shows as a whole-number error."""
    rng = np.random.default_rng(5)
    ev = gb.Eval(ctx)
    for (M, N, K) in [(999, 1001, 515), (2049, 260, 1030), (70, 70, 3500)]:
        for t1 in (False, True):
            for t2 in (False, True):
                sa, sb = ((K, M) if t1 else (M, K)), ((N, K) if t2 else (K, N))
                a = rng.integers(-4, 5, sa).astype(np.float32)
                b = rng.integers(-4, 5, sb).astype(np.float32)
                got = to_np(ev.matmul(ev.constant(a), ev.constant(b), gb.MatProp(t1), gb.MatProp(t2)))[:, :, 0, 0]
                want = (a.T if t1 else a).astype(np.float64) @ (b.T if t2 else b).astype(np.float64)
                assert got.shape == (M, N)
                assert np.array_equal(got.astype(np.float64), want), (M, N, K, t1, t2)


def test_gemm_tc_ragged_mlp_gradients(ctx):
    """A tanh layer with batch 1000 and widths 520 -> 392: forward, dA and dW all ragged; dW accumulates (w used
    twice); gradients against f64."""
    rng = np.random.default_rng(9)
    B, D, H = 1000, 520, 392
    x = rng.standard_normal((B, D)).astype(np.float32)
    w = (rng.standard_normal((D, H)) / np.sqrt(D)).astype(np.float32)
    g = gb.Graph1(ctx)
    vx, vw = g.variable(x), g.variable(w)
    y = g.add(g.tanh(g.matmul_nn(vx, vw)), g.matmul_nn(vx, vw))
    loss = g.norm2(y)
    store = g.evaluate_gradients_once(loss.gid(), 1.0)
    X, W = x.astype(np.float64), w.astype(np.float64)
    Z = X @ W
    Y = np.tanh(Z) + Z
    dZ = 2 * Y * (1 - np.tanh(Z) ** 2) + 2 * Y
    for got, want in [(store.get(vw.gid()), X.T @ dZ), (store.get(vx.gid()), dZ @ W.T)]:
        got = to_np(got)[:, :, 0, 0].astype(np.float64)
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()


@pytest.mark.parametrize("ad,bd", [((256, 128, 3), (128, 256)), ((256, 128), (128, 256, 2, 2)), ((300, 200, 2), (200, 260, 2))])
def test_batched_matmul_takes_the_tensor_core_path(ctx, ad, bd):
    """af::matmul over batch dims 2, 3 with broadcast (gad/src/matrix.rs:108-114): one tensor-core launch per matrix of the
    batch, against f64 numpy under the 3xTF32 bound."""
    rng = np.random.default_rng(17)
    a, b = rng.standard_normal(ad).astype(np.float32), rng.standard_normal(bd).astype(np.float32)
    ev = gb.Eval(ctx)
    ctx.profile_begin()
    got = to_np(ev.matmul(ev.constant(a), ev.constant(b)))
    prof = ctx.profile_end()
    a4, b4 = a.reshape(ad + (1,) * (4 - len(ad))), b.reshape(bd + (1,) * (4 - len(bd)))
    batch = (max(a4.shape[2], b4.shape[2]), max(a4.shape[3], b4.shape[3]))
    assert got.shape == (ad[0], bd[1]) + batch
    K = ad[1]
    for i in range(batch[0]):
        for j in range(batch[1]):
            A = a4[:, :, i % a4.shape[2], j % a4.shape[3]].astype(np.float64)
            B = b4[:, :, i % b4.shape[2], j % b4.shape[3]].astype(np.float64)
            bound = (2.0 ** -20 + 3 * (K / 8) * 2.0 ** -23) * (np.abs(A) @ np.abs(B))
            assert np.all(np.abs(got[:, :, i, j].astype(np.float64) - A @ B) <= bound), (i, j)
    assert prof["matmul"]["launches"] == 1  # one dispatch ...
