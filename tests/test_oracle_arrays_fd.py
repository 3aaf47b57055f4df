"""CPU: the oracle's ARRAY path checked the way the reference checks its own (gad/tests/*.rs `af` modules, helpers
gad/src/arrayfire.rs:114-164): one op at a time on small random arrays, the tape's gradient against central finite
differences of the forward evaluation, L-infinity < 1e-3 (2e-3 for matmul). The reference has no golden vectors for array
arithmetic (it lives in the un-vendored arrayfire crate), so this — the reference's own acceptance criterion — and the
known-answer tests of test_oracle_golden.py are what pin the checker that every GPU parity test compares against.
Differences are taken in f64 (the oracle's HostArray<f64>), where 1e-3 is a bound on the FORMULA, not on rounding."""
import numpy as np
import pytest


def estimate_gradient(oracle, x, direction, eps, f):
    """arrayfire.rs:120-150 over the oracle's Eval."""
    ev = oracle.OEval()
    v = np.array(x, dtype=np.float64)
    grad = np.zeros_like(v)
    flat, gflat = v.reshape(-1), grad.reshape(-1)
    d = np.asarray(direction, dtype=np.float64)

    def y():
        out = np.asarray(f(ev, ev.constant(v.copy())).data(), dtype=np.float64)
        return float(np.sum(out.reshape(d.shape) * d))

    for i in range(flat.size):
        x0 = flat[i]
        flat[i] = x0 + eps
        y2 = y()
        flat[i] = x0 - eps
        y1 = y()
        gflat[i] = (y2 - y1) / (2 * eps)
        flat[i] = x0
    return grad


def tape_gradient(oracle, x, direction, f):
    g = oracle.OGraph1()
    a = g.variable(np.array(x, dtype=np.float64))
    b = f(g, a)
    store = g.evaluate_gradients_once(b.gid(), np.asarray(direction, dtype=np.float64))
    got = store.get(a.gid(), like=a)
    return None if got is None else np.asarray(got.data(), dtype=np.float64)


def rand(shape, seed, lo=-1.0, hi=1.0):
    return np.random.default_rng(seed).uniform(lo, hi, shape)


UNARY = [  # (name, f, input shape, output shape, input range)
    ("neg", lambda g, v: g.neg(v), (4, 3), (4, 3), (-1, 1)),
    ("exp", lambda g, v: g.exp(v), (4, 3), (4, 3), (-1, 1)),
    ("log", lambda g, v: g.log(v), (4, 3), (4, 3), (0.5, 2)),
    ("log1p", lambda g, v: g.log1p(v), (4, 3), (4, 3), (-0.4, 2)),
    ("sin", lambda g, v: g.sin(v), (4, 3), (4, 3), (-2, 2)),
    ("cos", lambda g, v: g.cos(v), (4, 3), (4, 3), (-2, 2)),
    ("tanh", lambda g, v: g.tanh(v), (4, 3), (4, 3), (-1, 1)),
    ("sigmoid", lambda g, v: g.sigmoid(v), (4, 3), (4, 3), (-2, 2)),
    ("reciprocal", lambda g, v: g.reciprocal(v), (4, 3), (4, 3), (0.5, 2)),
    ("sqrt", lambda g, v: g.sqrt(v), (4, 3), (4, 3), (0.5, 2)),
    ("powc4", lambda g, v: g.powc(v, 4), (4, 3), (4, 3), (-1, 1)),
    ("mulc", lambda g, v: g.mulc(v, 3), (4, 3), (4, 3), (-1, 1)),
    ("addc", lambda g, v: g.addc(v, 3), (4, 3), (4, 3), (-1, 1)),
    ("abs", lambda g, v: g.abs(v), (4, 3), (4, 3), (0.2, 1)),
    ("relu", lambda g, v: g.relu(v), (4, 3), (4, 3), (0.2, 1)),
    ("square", lambda g, v: g.mul(v, v), (4, 3), (4, 3), (-1, 1)),
    ("flat", lambda g, v: g.flat(v), (4, 3), (12,), (-1, 1)),
    ("moddims", lambda g, v: g.moddims(v, (2, 2, 3)), (4, 3), (2, 2, 3), (-1, 1)),
    ("sum_as", lambda g, v: g.sum_as(v, (1, 3)), (4, 3), (1, 3), (-1, 1)),
    ("sum_as_all", lambda g, v: g.sum_as(v, (1,)), (4, 3), (1,), (-1, 1)),
    ("tile_as", lambda g, v: g.tile_as(v, (4, 3)), (1, 3), (4, 3), (-1, 1)),
    ("transpose", lambda g, v: g.transpose(v), (4, 3), (3, 4), (-1, 1)),
    ("max_as", lambda g, v: g.max_as(v, (1, 3)), (4, 3), (1, 3), (-1, 1)),
    ("softmax_as", lambda g, v: g.softmax_as(v, (1, 3)), (4, 3), (4, 3), (-1, 1)),
    ("tanh_chain", lambda g, v: g.tanh(g.add(g.mul(v, v), g.exp(v))), (4, 3), (4, 3), (-1, 1)),
]


@pytest.mark.parametrize("name,f,shape,out_shape,rng", UNARY, ids=[c[0] for c in UNARY])
def test_oracle_vjp_matches_finite_differences(oracle, name, f, shape, out_shape, rng):
    x = rand(shape, 31, *rng)
    # a non-constant direction: a wrong transpose / tiling in a VJP cannot hide behind all-ones
    direction = rand(out_shape, 32, 0.5, 1.5)
    got = tape_gradient(oracle, x, direction, f)
    est = estimate_gradient(oracle, x, direction, 1e-3, f)
    assert got is not None and got.shape[:len(shape)] == tuple(shape) or got.size == x.size
    assert np.max(np.abs(got.reshape(-1) - est.reshape(-1))) < 1e-3, name


@pytest.mark.parametrize("t1,t2", [(False, False), (True, False), (False, True), (True, True)])
def test_oracle_matmul_vjp_matches_finite_differences(oracle, t1, t2):
    """tests/matrix.rs:25-49 for the untransposed product. For transposed operands the reference's VJP formula
    (matrix.rs:164-176: dA = matmul(G, B, p1, p2^T), dB = matmul(A, G, p1^T, p2)) is the right adjoint only where it
    happens to type-check; elsewhere it returns Error::Dimensions. The oracle restates the formula AS IS
    (the device backend reproduces it under gadcu_ctx_set_strict_reference and computes the correct adjoint by default):
    so every transposed case must either agree with finite differences or fail with Dimensions — never a wrong value."""
    import gad_b200 as gb  # (MatProp only: a plain pair of flags)
    M, K, N = 4, 3, 5
    xa = rand((K, M) if t1 else (M, K), 33)
    xb = rand((N, K) if t2 else (K, N), 34)
    direction = rand((M, N), 35, 0.5, 1.5)
    p1, p2 = gb.MatProp(t1), gb.MatProp(t2)
    fa = lambda g, v: g.matmul(v, g.constant(xb), p1, p2)
    fb = lambda g, v: g.matmul(g.constant(xa), v, p1, p2)
    agreed = 0
    for x, f in [(xa, fa), (xb, fb)]:
        est = estimate_gradient(oracle, x, direction, 1e-3, f)
        try:
            got = tape_gradient(oracle, x, direction, f)
        except oracle.OracleError as e:
            assert e.kind == "Dimensions" or getattr(e, "status", None) == 1, e
            assert t1 or t2, "the untransposed product must differentiate"
            continue
        assert got.shape == x.shape or got.size == x.size
        assert np.max(np.abs(got.reshape(-1) - est.reshape(-1))) < 2e-3, (t1, t2)
        agreed += 1
    if not (t1 or t2):
        assert agreed == 2


def test_oracle_binary_vjps_match_finite_differences(oracle):
    x0, y0 = rand((4, 3), 36, 0.5, 2.0), rand((4, 3), 37, 0.5, 2.0)
    direction = rand((4, 3), 38, 0.5, 1.5)
    for name in ["add", "sub", "mul", "div", "pow", "min", "max"]:
        fa = lambda g, v, name=name: getattr(g, name)(v, g.constant(y0))
        fb = lambda g, v, name=name: getattr(g, name)(g.constant(x0), v)
        for x, f in [(x0, fa), (y0, fb)]:
            got, est = tape_gradient(oracle, x, direction, f), estimate_gradient(oracle, x, direction, 1e-3, f)
            if got is None:  # select_argmax leaves the unselected side ABSENT (tests/compare.rs:90-91): its derivative is zero
                got = np.zeros_like(est)
            assert np.max(np.abs(got.reshape(-1) - est.reshape(-1))) < 1e-3, name


def test_oracle_mlp_loss_gradient_matches_finite_differences(oracle):
    """The C4 tape at toy size: tanh(x.W1 + tile(b1)) . W2, square loss; dL/dW1 and dL/db1 against central differences."""
    B, D, H = 5, 4, 3
    x, t = rand((B, D), 40), rand((B, H), 41)
    w1, b1, w2 = rand((D, D), 42), rand((1, D), 43), rand((D, H), 44)

    def loss(g, vw1, vb1):
        h = g.tanh(g.add(g.matmul_nn(g.constant(x), vw1), g.tile_as(vb1, (B, D))))
        out = g.matmul_nn(h, g.constant(w2))
        delta = g.sub(g.constant(t), out)
        return g.sum_as(g.mul(delta, delta), (1,))

    one = np.ones((1,))
    got_w = tape_gradient(oracle, w1, one, lambda g, v: loss(g, v, g.constant(b1)))
    est_w = estimate_gradient(oracle, w1, one, 1e-4, lambda g, v: loss(g, v, g.constant(b1)))
    got_b = tape_gradient(oracle, b1, one, lambda g, v: loss(g, g.constant(w1), v))
    est_b = estimate_gradient(oracle, b1, one, 1e-4, lambda g, v: loss(g, g.constant(w1), v))
    assert np.max(np.abs(got_w.reshape(-1) - est_w.reshape(-1))) < 1e-5
    assert np.max(np.abs(got_b.reshape(-1) - est_b.reshape(-1))) < 1e-5
