"""GPU parity: every operator of the AfAlgebra bundle, forward and reverse sweep, CUDA backend vs the
CPU oracle on the same seeded inputs — through the C ABI.

Tolerances (stated per SURVEY §8c, confirmed on B200):
  * dims, ids, present/absent store entries, error variants: exact.
  * +, -, *, /, sqrt, neg, select, tile, moddims, transpose: bit-exact (IEEE, no FMA contraction).
  * exp/log/log1p/sin/cos/tanh/pow and VJPs through them: <= 4 ulp f32, <= 4 ulp f64
    (CUDA libm vs glibc, each <= 2 ulp from the true value).
  * reductions (sum_as, dot): |err| <= 8 * eps * sum|x_i| (tree order on GPU vs sequential on CPU).
  * matmul: |err| <= 2^-20 * sum_k |a_k||b_k| for f32 (covers 3xTF32), 2^-48 for f64.
  * fused vs unfused reverse sweep on the GPU: bit-exact.
"""
import numpy as np
import pytest

import gad_b200 as gb
from helpers import rand, to_np, ulp_diff

pytestmark = pytest.mark.gpu

DT = {"f32": np.float32, "f64": np.float64}
EXACT_UNARY = ["neg", "reciprocal", "sqrt", "zeros", "ones", "abs", "relu", "sign"]
LIBM_UNARY = ["exp", "log", "log1p", "sin", "cos", "tanh", "sigmoid"]
SIZES = [(4, 3), (1,), (5,), (1023,), (4097, 3), (2, 2, 3), (3, 1, 2, 2)]


def run_both(ctx, oracle, build, inputs, seed_fn, once=True, fuse=True):
    """build(g, vars) -> output value. Returns dict with forward + gradients from both backends."""
    ctx.set_fusion(fuse)
    g = gb.Graph1(ctx)
    og = oracle.OGraph1()
    vs = [g.variable(x) for x in inputs]
    ovs = [og.variable(x) for x in inputs]
    out, oout = build(g, vs), build(og, ovs)
    fwd, ofwd = to_np(out), to_np(oout)
    seed = seed_fn(ofwd)
    if out.id() is None:
        assert oout.id() is None
        return dict(fwd=fwd, ofwd=ofwd, grads=[None] * len(vs), ograds=[None] * len(vs))
    ev = g.evaluate_gradients_once if once else g.evaluate_gradients
    oev = og.evaluate_gradients_once if once else og.evaluate_gradients
    store, ostore = ev(out.gid(), seed), oev(oout.gid(), seed)
    grads, ograds = [], []
    for v, ov in zip(vs, ovs):
        a, b = store.get(v.gid()), ostore.get(ov.gid(), like=ov)
        grads.append(None if a is None else to_np(a))
        ograds.append(None if b is None else to_np(b))
    assert sorted(store.ids()) == sorted(ostore.ids())
    ctx.set_fusion(True)
    return dict(fwd=fwd, ofwd=ofwd, grads=grads, ograds=ograds)


def assert_close(a, b, ulps):
    if a is None or b is None:
        assert a is None and b is None
        return
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (a.shape, b.shape, a.dtype, b.dtype)
    if ulps == 0:
        assert np.array_equal(a, b, equal_nan=True), f"max ulp diff {ulp_diff(a, b)}"
    else:
        assert ulp_diff(a, b) <= ulps, f"max ulp diff {ulp_diff(a, b)} > {ulps}"


def ones_like_seed(ofwd):
    return np.ones_like(ofwd) if isinstance(ofwd, np.ndarray) else type(ofwd)(1.0)


def direction_seed(ofwd):
    if not isinstance(ofwd, np.ndarray):
        return type(ofwd)(1.7)
    return (0.5 + np.arange(ofwd.size, dtype=ofwd.dtype).reshape(ofwd.shape, order="F") % 7) / 3


@pytest.mark.parametrize("dt", ["f32", "f64"])
@pytest.mark.parametrize("shape", SIZES, ids=str)
@pytest.mark.parametrize("op", EXACT_UNARY + LIBM_UNARY)
def test_unary(ctx, oracle, op, shape, dt):
    x = rand(shape, 11, 0.25, 2.0, DT[dt])
    if op in ("abs", "relu", "sign", "neg"):
        x = x - 1.0
    r = run_both(ctx, oracle, lambda g, v: getattr(g, op)(v[0]), [x], direction_seed)
    ulps = 0 if op in EXACT_UNARY else 4
    assert_close(r["fwd"], r["ofwd"], ulps)
    # A VJP chains up to 5 rounded ops through a libm value. tanh's (1 - t^2) and sigmoid's c(1 - c)
    # cancel: a 4-ulp difference in t is amplified by 2t^2/(1-t^2) (x26 at x=2), so their budget is wider.
    budget = {"tanh": 512, "sigmoid": 128}.get(op, 16)
    assert_close(r["grads"][0], r["ograds"][0], 0 if op in EXACT_UNARY else budget)


@pytest.mark.parametrize("dt", ["f32", "f64"])
@pytest.mark.parametrize("shape", SIZES, ids=str)
@pytest.mark.parametrize("op", ["add", "sub", "mul", "div", "pow", "min", "max"])
def test_binary(ctx, oracle, op, shape, dt):
    a, b = rand(shape, 1, 0.5, 2.0, DT[dt]), rand(shape, 2, 0.5, 2.0, DT[dt])
    r = run_both(ctx, oracle, lambda g, v: getattr(g, op)(v[0], v[1]), [a, b], direction_seed)
    exact = op != "pow"
    assert_close(r["fwd"], r["ofwd"], 0 if exact else 4)
    for ga, gb_ in zip(r["grads"], r["ograds"]):
        assert_close(ga, gb_, 0 if exact else 32)


@pytest.mark.parametrize("dt", ["f32", "f64"])
@pytest.mark.parametrize("op,c", [("addc", 4.0), ("addc", 1), ("mulc", 4), ("mulc", 0.3), ("powc", 4), ("powc", 2.5), ("setc", 3)])
def test_const_arith(ctx, oracle, op, c, dt):
    x = rand((4, 3), 5, 0.5, 2.0, DT[dt])
    r = run_both(ctx, oracle, lambda g, v: getattr(g, op)(v[0], c), [x], direction_seed)
    assert_close(r["fwd"], r["ofwd"], 4 if op == "powc" else 0)
    assert_close(r["grads"][0], r["ograds"][0], 16 if op == "powc" else 0)


@pytest.mark.parametrize("variant", ["both", "r0", "r1", "none"])
def test_select_argmax(ctx, oracle, variant):
    a, b, c, d = (rand((4, 3), s) for s in (1, 2, 3, 4))

    def build(g, v):
        r0 = v[2] if variant in ("both", "r0") else None
        r1 = v[3] if variant in ("both", "r1") else None
        return g.select_argmax(v[0], v[1], r0, r1)

    r = run_both(ctx, oracle, build, [a, b, c, d], direction_seed)
    assert_close(r["fwd"], r["ofwd"], 0)
    for x, y in zip(r["grads"], r["ograds"]):
        assert_close(x, y, 0)
    assert r["grads"][0] is None and r["grads"][1] is None  # compared values get no gradient


@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_shape_ops(ctx, oracle, dt):
    x = rand((4, 3), 7, dtype=DT[dt])
    for build in [lambda g, v: g.flat(v[0]), lambda g, v: g.moddims(v[0], (2, 2, 3)),
                  lambda g, v: g.transpose(v[0]), lambda g, v: g.transpose(g.transpose(v[0]), True)]:
        r = run_both(ctx, oracle, build, [x], direction_seed)
        assert_close(r["fwd"], r["ofwd"], 0)
        assert_close(r["grads"][0], r["ograds"][0], 0)
    # tiling a non-singleton axis has no valid VJP in the reference (sum_as(g, vdims) rejects it,
    # array.rs:159-170): both sides must report Dimensions from the sweep
    g, og = gb.Graph1(ctx), oracle.OGraph1()
    v = rand((2, 1, 2), 8, dtype=DT[dt])
    t, ot = g.tile_as(g.variable(v), (4, 3, 2)), og.tile_as(og.variable(v), (4, 3, 2))
    assert np.array_equal(to_np(t), to_np(ot))
    with pytest.raises(gb.DimensionsError):
        g.evaluate_gradients_once(t.gid(), np.ones((4, 3, 2), DT[dt]))
    with pytest.raises(oracle.OracleError) as e:
        og.evaluate_gradients_once(ot.gid(), np.ones((4, 3, 2), DT[dt]))
    assert e.value.kind == "Dimensions"
    for vshape, rd in [((1, 3), (4, 3)), ((4, 1), (4, 3)), ((1,), (4, 3)), ((1, 1, 2), (4, 3, 2)), ((1, 300), (200, 300))]:
        v = rand(vshape, 8, dtype=DT[dt])
        r = run_both(ctx, oracle, lambda g, vv: g.tile_as(vv[0], rd), [v], direction_seed)
        assert_close(r["fwd"], r["ofwd"], 0)
        got, want = r["grads"][0], r["ograds"][0]
        assert got.shape == want.shape
        eps = np.finfo(DT[dt]).eps
        assert np.all(np.abs(got - want) <= 8 * eps * np.abs(want) * (np.prod(rd) / np.prod(vshape)) + 1e-30)


@pytest.mark.parametrize("dt", ["f32", "f64"])
@pytest.mark.parametrize("vshape,rd", [((4, 3), (1, 3)), ((4, 3), (4, 1)), ((4, 3), (1,)), ((5, 4, 3), (5, 1, 3)),
                                        ((1000, 37), (1, 37)), ((1000, 37), (1000, 1)), ((70001,), (1,)),
                                        ((3, 4, 2, 2), (1, 4, 1, 2)), ((8, 4), (8, 4))], ids=str)
def test_reductions(ctx, oracle, vshape, rd, dt):
    x = rand(vshape, 9, -1.0, 1.0, DT[dt])
    eps = np.finfo(DT[dt]).eps
    for op in ["sum_as", "max_as"]:
        r = run_both(ctx, oracle, lambda g, v: getattr(g, op)(v[0], rd), [x], direction_seed, once=False)
        if op == "max_as":
            assert_close(r["fwd"], r["ofwd"], 0)
        else:
            bound = 8 * eps * np.sum(np.abs(x)) / max(1, r["ofwd"].size) * (x.size / max(1, r["ofwd"].size)) + 8 * eps
            ref64 = np.asarray(x, dtype=np.float64)
            assert r["fwd"].shape == r["ofwd"].shape
            assert np.max(np.abs(r["fwd"].astype(np.float64) - r["ofwd"].astype(np.float64))) <= 8 * eps * np.sum(np.abs(ref64))
        assert_close(r["grads"][0], r["ograds"][0], 0 if op == "sum_as" else 4)


@pytest.mark.parametrize("dt", ["f32", "f64"])
def test_dot_norm2_scale_constant_as(ctx, oracle, dt):
    a, b = rand((4, 3), 1, -1, 1, DT[dt]), rand((4, 3), 2, -1, 1, DT[dt])
    eps = np.finfo(DT[dt]).eps
    r = run_both(ctx, oracle, lambda g, v: g.dot(v[0], v[1]), [a, b], ones_like_seed)
    assert abs(float(r["fwd"]) - float(r["ofwd"])) <= 8 * eps * np.sum(np.abs(a * b))
    assert_close(r["grads"][0], r["ograds"][0], 0)
    assert_close(r["grads"][1], r["ograds"][1], 0)
    # norm2 = dot(v, v): two contributions scale(g, v) accumulated as current + new (array.rs:42-44)
    r = run_both(ctx, oracle, lambda g, v: g.norm2(v[0]), [a], ones_like_seed)
    assert_close(r["grads"][0], r["ograds"][0], 0)
    # tests/array.rs:70-86: constant_as(norm2(a), [1])
    r = run_both(ctx, oracle, lambda g, v: g.constant_as(g.norm2(v[0]), (1,)), [a], ones_like_seed)
    assert_close(r["grads"][0], r["ograds"][0], 0)
    # scale(lambda, v) with a tracked scalar: lambda gets dot(g, v), v gets scale(lambda, g)
    r = run_both(ctx, oracle, lambda g, v: g.scale(g.as_scalar(g.sum_as(v[1], (1,))), v[0]), [a, b], direction_seed)
    assert np.allclose(r["fwd"], r["ofwd"], rtol=64 * eps, atol=64 * eps)
    assert np.allclose(r["grads"][0], r["ograds"][0], rtol=64 * eps, atol=64 * eps)
    assert np.allclose(r["grads"][1], r["ograds"][1], rtol=64 * eps, atol=64 * eps)


@pytest.mark.parametrize("dt", ["f32", "f64"])
@pytest.mark.parametrize("m,k,n", [(4, 3, 5), (2, 2, 2), (1, 1, 1), (65, 33, 17), (128, 64, 96), (100, 257, 3)])
def test_matmul_nn(ctx, oracle, m, k, n, dt):
    """tests/matrix.rs:25-49 shapes and ragged ones; the reference's VJP formulas are only valid for
    non-transposed operands, so gradients are compared for NN."""
    a, b = rand((m, k), 3, -1, 1, DT[dt]), rand((k, n), 4, -1, 1, DT[dt])
    r = run_both(ctx, oracle, lambda g, v: g.matmul_nn(v[0], v[1]), [a, b], direction_seed)
    scale = 2.0 ** -20 if dt == "f32" else 2.0 ** -48
    a64, b64 = np.abs(a.astype(np.float64)), np.abs(b.astype(np.float64))
    bound = scale * (a64 @ b64)
    assert np.all(np.abs(r["fwd"][:, :, 0, 0].astype(np.float64) - r["ofwd"][:, :, 0, 0].astype(np.float64)) <= bound + 1e-300)
    seed = direction_seed(r["ofwd"])[:, :, 0, 0].astype(np.float64)
    ga_bound = scale * (np.abs(seed) @ b64.T)
    gb_bound = scale * (a64.T @ np.abs(seed))
    assert np.all(np.abs(r["grads"][0][:, :, 0, 0].astype(np.float64) - r["ograds"][0][:, :, 0, 0]) <= ga_bound + 1e-300)
    assert np.all(np.abs(r["grads"][1][:, :, 0, 0].astype(np.float64) - r["ograds"][1][:, :, 0, 0]) <= gb_bound + 1e-300)


def test_matmul_transposed_and_batched_forward(ctx, oracle):
    ev, oev = gb.Eval(ctx), oracle.OEval()
    for (ad, bd, t1, t2) in [((3, 4), (3, 5), True, False), ((4, 3), (5, 3), False, True), ((3, 4), (5, 3), True, True),
                             ((4, 3, 2, 2), (3, 5), False, False), ((4, 3), (3, 5, 3), False, False),
                             ((4, 3, 2), (3, 5, 2), False, False)]:
        a, b = rand(ad, 1, -1, 1), rand(bd, 2, -1, 1)
        got = ev.matmul(ev.constant(a), ev.constant(b), gb.MatProp(t1), gb.MatProp(t2))
        want = oev.matmul(oev.constant(a), oev.constant(b), gb.MatProp(t1), gb.MatProp(t2))
        assert got.dims() == want.dims()
        assert np.allclose(to_np(got), to_np(want), rtol=1e-5, atol=1e-6)
    # the reference's matmul VJP with a transposed operand is dimensionally invalid: same error
    g, og = gb.Graph1(ctx), oracle.OGraph1()
    a, b = rand((3, 4), 1), rand((3, 5), 2)
    va, vb = g.variable(a), g.variable(b)
    c = g.matmul(va, vb, gb.MatProp(True), gb.MatProp())
    oc = og.matmul(og.variable(a), og.variable(b), gb.MatProp(True), gb.MatProp())
    ctx.set_strict_reference(True)
    try:
        with pytest.raises(gb.DimensionsError):
            g.evaluate_gradients(c.gid(), np.ones((4, 5), np.float32))
    finally:
        ctx.set_strict_reference(False)
    with pytest.raises(oracle.OracleError) as e:
        og.evaluate_gradients_once(oc.gid(), np.ones((4, 5), np.float32))
    assert e.value.kind == "Dimensions"
    # default mode: the correct adjoints (C = A^T B: dA = B G^T, dB = A G), against f64 numpy,
    # for every transposition pattern
    for t1, t2 in [(True, False), (False, True), (True, True), (False, False)]:
        sa, sb = ((3, 4) if t1 else (4, 3)), ((5, 3) if t2 else (3, 5))
        a, b = rand(sa, 5, -1, 1), rand(sb, 6, -1, 1)
        g = gb.Graph1(ctx)
        va, vb = g.variable(a), g.variable(b)
        c = g.matmul(va, vb, gb.MatProp(t1), gb.MatProp(t2))
        seed = rand((4, 5), 7, -1, 1)
        store = g.evaluate_gradients_once(c.gid(), seed)
        A, B, G = a.astype(np.float64), b.astype(np.float64), seed.astype(np.float64)
        opA, opB = (A.T if t1 else A), (B.T if t2 else B)
        dA = G @ opB.T
        dB = opA.T @ G
        dA, dB = (dA.T if t1 else dA), (dB.T if t2 else dB)
        assert np.allclose(to_np(store.get(va.gid()))[:, :, 0, 0], dA, rtol=1e-5, atol=1e-6), (t1, t2)
        assert np.allclose(to_np(store.get(vb.gid()))[:, :, 0, 0], dB, rtol=1e-5, atol=1e-6), (t1, t2)


@pytest.mark.parametrize("rd", [(1, 3), (4, 1), (1,)])
def test_array_compare(ctx, oracle, rd):
    """tests/array_compare.rs: argmax_as is a constant whose mass is 3 / 4 / 1; softmax likewise."""
    x = rand((4, 3), 12)
    g = gb.Graph1(ctx)
    a = g.variable(x)
    am = g.argmax_as(a, rd)
    assert am.id() is None
    mass = {(1, 3): 3.0, (4, 1): 4.0, (1,): 1.0}[rd]
    assert abs(g.as_scalar(g.sum_as(am, (1,))).data().item() - mass) < np.finfo(np.float32).eps * 4
    assert abs(g.as_scalar(g.sum_as(g.softmax_as(a, rd), (1,))).data().item() - mass) < 1e-6
    for op in ["softmax_as", "max_as"]:
        r = run_both(ctx, oracle, lambda gg, v: getattr(gg, op)(v[0], rd), [x], direction_seed, once=False)
        assert np.allclose(r["fwd"], r["ofwd"], rtol=1e-6, atol=1e-7)
        assert np.allclose(r["grads"][0], r["ograds"][0], rtol=1e-5, atol=1e-6)


CHAINS = {
    "tanh_mul": lambda g, v: g.mul(g.tanh(v[0]), v[1]),
    "accumulate_three": lambda g, v: g.add(g.add(g.mul(v[0], v[0]), g.exp(v[0])), g.mul(v[0], v[1])),
    "div_chain": lambda g, v: g.div(g.sqrt(v[0]), g.addc(g.sigmoid(v[1]), 1)),
    "sub_neg": lambda g, v: g.sub(g.neg(v[0]), g.mulc(g.reciprocal(v[1]), 3)),
    "log1p_cos": lambda g, v: g.mul(g.log1p(v[0]), g.cos(g.sin(v[1]))),
    "powc_pow": lambda g, v: g.add(g.powc(v[0], 3), g.pow(v[0], v[1])),
    "relu_abs": lambda g, v: g.add(g.relu(g.addc(v[0], -1.0)), g.abs(g.sub(v[0], v[1]))),
    "add_all": lambda g, v: g.add_all([v[0], g.mul(v[0], v[1]), v[1], v[0]]),
    "reduce_tile": lambda g, v: g.mul(g.tile_as(g.sum_as(v[0], (1, 3)), (4, 3)), v[1]),
    "dot_scale": lambda g, v: g.scale(g.dot(v[0], v[1]), g.tanh(v[0])),
}


@pytest.mark.parametrize("name", sorted(CHAINS))
def test_fused_sweep_is_bit_identical_to_the_reference_op_sequence(ctx, oracle, name):
    """The fused VJP+accumulate kernels compute, per element, exactly the rounded values of the
    reference's one-primitive-per-kernel sweep: compare the two GPU sweeps with ==, then the oracle."""
    a, b = rand((4, 3), 21, 0.5, 2.0), rand((4, 3), 22, 0.5, 2.0)
    fused = run_both(ctx, oracle, CHAINS[name], [a, b], direction_seed, fuse=True)
    plain = run_both(ctx, oracle, CHAINS[name], [a, b], direction_seed, fuse=False)
    for x, y in zip(fused["grads"], plain["grads"]):
        assert_close(x, y, 0)
    assert_close(fused["fwd"], plain["fwd"], 0)
    for x, y in zip(fused["grads"], fused["ograds"]):
        assert x.shape == y.shape
        assert np.allclose(x, y, rtol=2e-5, atol=1e-6)


def test_large_ragged_elementwise_matches_oracle(ctx, oracle):
    """Vector main loop + scalar tail + grid-stride, at a size that is not a multiple of anything."""
    n = 1_000_003
    x, y = rand((n,), 3, -1, 1), rand((n,), 4, 0.5, 2.0)
    r = run_both(ctx, oracle, lambda g, v: g.add(g.mul(g.exp(v[0]), g.log(v[1])), v[0]), [x, y], ones_like_seed)
    # exp(x)*log(y) + x and its x-derivative exp(x)*log(y) + 1 cancel, so the bound is in terms of the
    # magnitude of the summands: 4 ulp (libm) on the product plus the rounding of the sum
    eps = np.finfo(np.float32).eps
    prod = np.abs(np.exp(x.astype(np.float64)) * np.log(y.astype(np.float64)))
    assert np.all(np.abs(r["fwd"].reshape(-1).astype(np.float64) - r["ofwd"].reshape(-1)) <= 6 * eps * (prod + np.abs(x)))
    assert np.all(np.abs(r["grads"][0].reshape(-1).astype(np.float64) - r["ograds"][0].reshape(-1)) <= 6 * eps * (prod + 1.0))
    assert_close(r["grads"][1], r["ograds"][1], 8)


def test_errors_are_values_and_launch_nothing(ctx, oracle):
    """Dimension checks run on the host before any launch (arith.rs:66): a failing op adds no kernel."""
    g = gb.Graph1(ctx)
    a, b = g.variable(rand((4, 3), 1)), g.variable(rand((3, 4), 2))
    ctx.synchronize()
    before = ctx.launches()
    for fn in [lambda: g.add(a, b), lambda: g.mul(a, b), lambda: g.matmul_nn(a, a), lambda: g.sum_as(a, (2, 3)),
               lambda: g.as_scalar(a), lambda: g.dot(a, b), lambda: g.moddims(a, (5, 5)), lambda: g.tile_as(a, (6, 3))]:
        with pytest.raises(gb.DimensionsError):
            fn()
    with pytest.raises(gb.ReducedDimensionsError):
        g.max_as(a, (2, 3))
    assert ctx.launches() == before
    nodes = g.num_nodes()
    assert nodes == 2  # failed ops record no node
    # sweep errors: wrong seed dims (graph.rs:156), constant root (MissingId), foreign id (MissingNode)
    c = g.neg(a)
    with pytest.raises(gb.DimensionsError):
        g.evaluate_gradients(c.gid(), np.ones((3, 4), np.float32))
    with pytest.raises(gb.MissingIdError):
        g.constant(rand((2,), 1)).gid()
    g2 = gb.Graph1(ctx)
    g2.variable(rand((4, 3), 1))
    with pytest.raises(gb.MissingNodeError):
        g2.evaluate_gradients(c.gid(), np.ones((4, 3), np.float32))
    with pytest.raises(gb.TypeMismatchError):
        g.add(a, g.variable(rand((4, 3), 1, dtype=np.float64)))
