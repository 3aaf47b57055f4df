"""GPU: parity at the sizes bench.py runs, against f64 — what tests/test_gpu_full_size.py's self-consistency properties
cannot show.

  * the three GEMM shapes of the C4 step at B = 8192, D = 4096 — forward NN, dA = G.W^T (NT), dW = h^T.G (TN, K = 8192:
    the shape whose ragged last wave takes the stream-K hand-over) — on 64 sampled rows x 64 sampled columns against
    f64 numpy, under the bound of tests/test_gpu_gemm_tc.py;
  * the real C4 step at full size: sampled dL/dW_l and all of dL/db_l against an f64 forward + backward in numpy;
  * C5 at B = 512, D = 256 (the tcgen05 path) against the f64 R-operator closed form (tests/mlp_f64.py), and at a toy
    size against the oracle's GraphN with the corrected matmul adjoint;
  * both matmul-VJP modes of the library against the matching oracle mode;
  * tanh / sigmoid VJPs against f64 with a bound that follows the cancellation in 1 - t^2 / c (1 - c) instead of a wide
    ulp budget.
"""
import numpy as np
import pytest

import gad_b200 as gb
from gad_b200 import workloads as wl
from helpers import rand, to_np
import mlp_f64

pytestmark = pytest.mark.gpu

B_FULL, D_FULL = 8192, 4096


def _gemm_bound(absA_rows, absB_cols, K):
    S = absA_rows @ absB_cols
    return (2.0 ** -20 + 3 * (K / 8) * 2.0 ** -23) * S, S


@pytest.mark.parametrize("name,M,N,K,t1,t2", [("forward NN", B_FULL, D_FULL, D_FULL, False, False),
                                               ("dA = G.W^T (NT)", B_FULL, D_FULL, D_FULL, False, True),
                                               ("dW = h^T.G (TN, stream-K)", D_FULL, D_FULL, B_FULL, True, False)])
def test_bench_gemm_shapes_sampled_against_f64(ctx, name, M, N, K, t1, t2):
    sa, sb = ((K, M) if t1 else (M, K)), ((N, K) if t2 else (K, N))
    a = ctx.random(sa, 21, normal=True, a=0.0, b=1.0)
    b = ctx.random(sb, 22, normal=True, a=0.0, b=1.0)
    ev = gb.Eval(ctx)
    got = to_np(ev.matmul(ev.constant(a), ev.constant(b), gb.MatProp(t1), gb.MatProp(t2)))[:, :, 0, 0]
    assert got.shape == (M, N)
    ha, hb = to_np(a)[:, :, 0, 0], to_np(b)[:, :, 0, 0]
    rng = np.random.default_rng(7)
    # sampled rows / columns: some from every 256 x 256 tile class (first, last, last-wave tiles) plus random ones
    rows = np.unique(np.concatenate([[0, 127, 128, 255, 256, M - 256, M - 129, M - 1], rng.choice(M, 56, replace=False)]))
    cols = np.unique(np.concatenate([[0, 127, 128, 255, 256, N - 256, N - 129, N - 1], rng.choice(N, 56, replace=False)]))
    A = (ha[:, rows].T if t1 else ha[rows, :]).astype(np.float64)   # [r, K]
    Bm = (hb[cols, :].T if t2 else hb[:, cols]).astype(np.float64)  # [K, c]
    want = A @ Bm
    bound, S = _gemm_bound(np.abs(A), np.abs(Bm), K)
    err = np.abs(got[np.ix_(rows, cols)].astype(np.float64) - want)
    assert np.all(err <= bound), (name, float(err.max()), float((err / bound).max()))
    assert float(np.sqrt(np.mean((err / S) ** 2))) <= 2.0 ** -20, name


def test_c4_full_size_sampled_gradients_against_f64(ctx):
    B, D, L = B_FULL, D_FULL, 3
    X = ctx.random((B, D), 5, normal=True, a=0.0, b=1.0)
    T = ctx.random((B, D), 9, normal=True, a=0.0, b=1.0)
    Ws = [ctx.random((D, D), 6 + l, normal=True, a=0.0, b=1.0 / np.sqrt(D)) for l in range(L)]
    bs = [ctx.array(np.full((1, D), 0.01, np.float32)) for _ in range(L)]
    loss, gW, gb_ = wl.mlp_step(ctx, X, Ws, bs, T)
    hX, hT = to_np(X)[:, :, 0, 0], to_np(T)[:, :, 0, 0]
    hW, hb = [to_np(w)[:, :, 0, 0] for w in Ws], [to_np(b)[:, :, 0, 0] for b in bs]
    acts = mlp_f64.mlp_forward_f64(hX, hW, hb)           # 3 dgemm of 8192 x 4096 x 4096
    deltas, want_loss = mlp_f64.mlp_deltas_f64(acts, hW, hT)  # 2 more
    # the loss is a dot product over 33.5 M f32 squares of residuals that each carry the forward pass's f32 rounding
    # (three K = 4096 products deep): measured 1.8e-5 of the f64 value on B200
    assert abs(loss.item() - want_loss) <= 1e-4 * abs(want_loss)
    rng = np.random.default_rng(11)
    rows, cols = np.sort(rng.choice(D, 64, replace=False)), np.sort(rng.choice(D, 64, replace=False))
    eps = np.finfo(np.float32).eps
    for l in range(L):
        h, d = acts[l], deltas[l]
        want = h[:, rows].T @ d[:, cols]
        got = to_np(gW[l])[:, :, 0, 0][np.ix_(rows, cols)].astype(np.float64)
        # the GEMM bound on the last product (its operands carry the f32 rounding of everything upstream, which the
        # second term covers: each h, delta element is within a few eps (relative to its layer's scale) of the f64 one,
        # and those errors add up like a random walk over the B summands)
        S = np.abs(h[:, rows]).T @ np.abs(d[:, cols])
        bound = (2.0 ** -20 + 3 * (B / 8) * 2.0 ** -23) * S + 64 * eps * np.sqrt(B) * np.max(np.abs(h)) * np.max(np.abs(d))
        err = np.abs(got - want)
        assert np.all(err <= bound), (l, float(err.max()), float((err / bound).max()))
        assert float(err.max()) <= 2e-5 * float(np.abs(want).max()) * np.sqrt(B) / 8, (l, float(err.max()), float(np.abs(want).max()))
        wb = d.sum(axis=0, keepdims=True)
        gbv = to_np(gb_[l])[:, :, 0, 0].astype(np.float64)
        sb = np.abs(d).sum(axis=0, keepdims=True)
        # the reduction's own bound plus the f32 rounding the delta elements carry from upstream (a random walk over B rows)
        assert np.all(np.abs(gbv - wb) <= 8 * eps * sb + 256 * eps * np.sqrt(B) * np.max(np.abs(d))), l


def test_c5_hvp_matches_f64_closed_form_at_tensor_core_shape(ctx):
    """B = 512, D = 256: every product of both sweeps takes the tcgen05 path (M, N >= 64, M.N.K >= 2^24)."""
    B, D = 512, 256
    X, Ws, bs, T = mlp_f64.mlp_inputs(B, D)
    rng = np.random.default_rng(10)
    Vs = [rng.standard_normal((D, D)).astype(np.float32) for _ in range(3)]
    launches = ctx.launches()
    ctx.profile_begin()
    hv, loss, nodes = wl.mlp_hvp(ctx, ctx.array(X), [ctx.array(w) for w in Ws], [ctx.array(b) for b in bs], ctx.array(T),
                                 [ctx.array(v) for v in Vs])
    prof = ctx.profile_end()
    # (22 products are recorded; the two no requested value depends on are never launched when products are deferred — at
    # this shape, too small for deferral, all 22 run)
    assert prof["matmul"]["launches"] >= 20 and ctx.launches() > launches
    want = mlp_f64.mlp_hvp_f64(X, Ws, bs, T, Vs)
    wl_, _, _ = mlp_f64.mlp_grad_f64(X, Ws, bs, T)
    assert abs(loss.item() - wl_) <= 1e-5 * abs(wl_)
    for l in range(3):
        got = to_np(hv[l])[:, :, 0, 0].astype(np.float64)
        rel = float(np.abs(got - want[l]).max() / np.abs(want[l]).max())
        assert rel <= 1e-4, (l, rel)


def test_c5_hvp_matches_the_oracle_with_corrected_adjoint(ctx, oracle):
    """The reference's own matmul VJP cannot produce C5 for B != D (matrix.rs:164-176 is a Dimensions error for the
    transposed operands of the recorded backward GEMMs); the library's default mode and the oracle's corrected mode
    compute the true adjoint there. Same tape, same op order: agreement to f32 rounding of the sums."""
    B, D = 64, 32
    X, Ws, bs, T = mlp_f64.mlp_inputs(B, D)
    rng = np.random.default_rng(10)
    Vs = [rng.standard_normal((D, D)).astype(np.float32) for _ in range(3)]
    hv, loss, nodes = wl.mlp_hvp(ctx, ctx.array(X), [ctx.array(w) for w in Ws], [ctx.array(b) for b in bs], ctx.array(T),
                                 [ctx.array(v) for v in Vs])
    oloss, ohv, _, onodes = oracle.mlp_hvp(X, Ws, bs, T, Vs, nthreads=2)
    assert nodes == onodes  # the recorded backward graphs have the same size
    assert abs(loss.item() - oloss) <= 1e-5 * abs(oloss)
    for l in range(3):
        got = to_np(hv[l])[:, :, 0, 0]
        assert np.abs(got - ohv[l]).max() <= 2e-5 * np.abs(ohv[l]).max(), l
    # and the oracle in the reference's own mode raises what the reference would
    with pytest.raises(oracle.OracleError) as e:
        og = oracle.OGraphN()
        a, b = og.variable(rand((3, 4), 1)), og.variable(rand((3, 5), 2))
        c = og.matmul(a, b, gb.MatProp(True), gb.MatProp())
        og.compute_gradients(c.gid(), og.constant(np.ones((4, 5), np.float32)))
    assert e.value.kind == "Dimensions"


@pytest.mark.parametrize("t1,t2", [(True, False), (False, True), (True, True), (False, False)])
def test_matmul_vjp_modes_against_the_matching_oracle_mode(ctx, oracle, t1, t2):
    """Default library mode == oracle with the corrected adjoint; strict mode == oracle as the reference wrote it
    (square operands, where the reference formula yields values rather than an error)."""
    n = 5
    a, b, seed = rand((n, n), 5, -1, 1), rand((n, n), 6, -1, 1), rand((n, n), 7, -1, 1)
    for strict in (False, True):
        before = oracle.set_corrected_adjoint(not strict)
        ctx.set_strict_reference(strict)
        try:
            g, og = gb.Graph1(ctx), oracle.OGraph1()
            va, vb = g.variable(a), g.variable(b)
            ova, ovb = og.variable(a), og.variable(b)
            c = g.matmul(va, vb, gb.MatProp(t1), gb.MatProp(t2))
            oc = og.matmul(ova, ovb, gb.MatProp(t1), gb.MatProp(t2))
            store, ostore = g.evaluate_gradients_once(c.gid(), seed), og.evaluate_gradients_once(oc.gid(), seed)
            for v, ov in ((va, ova), (vb, ovb)):
                got, want = to_np(store.get(v.gid())), to_np(ostore.get(ov.gid(), like=ov))
                assert np.allclose(got, want, rtol=1e-5, atol=1e-6), (t1, t2, strict)
        finally:
            ctx.set_strict_reference(False)
            oracle.set_corrected_adjoint(before)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("op", ["tanh", "sigmoid"])
def test_cancelling_vjps_against_f64_with_a_cancellation_aware_bound(ctx, op, dt):
    """d tanh = g (1 - t^2), d sigmoid = g c (1 - c). With t (or c) good to k ulp, 1 - t^2 carries an ABSOLUTE error of
    about 2 k eps t^2, which is large relative to 1 - t^2 when |t| -> 1: an ulp budget on the result has to be hundreds
    wide there and says nothing elsewhere. The bound below follows the error where it arises:
        |err| <= |g| eps (6 t^2 + 3 (1 - t^2))      tanh      (CUDA tanhf / tanh: <= 2 ulp)
        |err| <= |g| eps (6 c^2 + 4 c (1 - c))      sigmoid   (expf / exp: <= 2 ulp, then 1 + e, 1 / x, 1 - c, two products)
    """
    x = rand((4097, 3), 11, -6.0, 6.0, dt)
    seed = rand((4097, 3), 12, 0.5, 2.0, dt)
    g = gb.Graph1(ctx)
    v = g.variable(x)
    out = getattr(g, op)(v)
    store = g.evaluate_gradients_once(out.gid(), seed)
    got = to_np(store.get(v.gid()))[:, :, 0, 0].astype(np.float64)
    X, G = x.astype(np.float64), seed.astype(np.float64)
    eps = float(np.finfo(dt).eps)
    if op == "tanh":
        t = np.tanh(X)
        want, bound = G * (1 - t * t), np.abs(G) * eps * (6 * t * t + 3 * (1 - t * t))
    else:
        c = 1.0 / (1.0 + np.exp(-X))
        want, bound = G * c * (1 - c), np.abs(G) * eps * (6 * c * c + 4 * c * (1 - c))
    err = np.abs(got - want)
    assert np.all(err <= bound), (op, float((err / bound).max()))
