"""GPU: the reference's own tests replayed on the CUDA backend (gad/tests/*.rs, arrayfire modules),
plus the golden known answers through device scalars."""
import json
import os

import numpy as np
import pytest

import gad_b200 as gb
import programs
from helpers import rand, to_np

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "reference_known_answers.json")) as f:
    GOLDEN = json.load(f)


from gad_b200.testing import assert_almost_all_equal as almost_all_equal, estimate_gradient  # arrayfire.rs:114-164


@pytest.mark.parametrize("case", GOLDEN["scalar_f32_direction_1p7"], ids=lambda c: c["name"])
def test_scalar_analytic_known_answers_on_device(ctx, case):
    """tests/analytic.rs:11-169 with device scalars (1-element tapes)."""
    g = gb.Graph1(ctx)
    names = ["a", "b"]
    vs = {n: g.variable(ctx.scalar(v)) for n, v in zip(names, case["inputs"])}
    out = programs.run_expr(g, case["program"], **vs)
    assert abs(out.data().item() - case["value"]) < case["tol"]
    store = g.evaluate_gradients_once(out.gid(), 1.7)
    for n, want in zip(names, case["grads"]):
        assert abs(store.get(vs[n].gid()).item() - want) < case["tol"]


def test_gradient_simple_on_arrays(ctx):
    """tests/graph.rs:76-105: c = a + a.a (matmul), clone the tape, seeds 1 and 2, vs finite differences."""
    g = gb.Graph1(ctx)
    x = rand((2, 2), 1)
    a = g.variable(x)
    c = g.add(a, g.matmul_nn(a, a))
    est = estimate_gradient(ctx, x, np.ones((2, 2), np.float32), 0.001, lambda ev, v: ev.add(ev.matmul_nn(v, v), v))
    g1 = g.clone().evaluate_gradients_once(c.gid(), np.ones((2, 2), np.float32))
    g2 = g.evaluate_gradients_once(c.gid(), np.full((2, 2), 2.0, np.float32))
    ga1, ga2 = to_np(g1.get(a.gid())), 0.5 * to_np(g2.get(a.gid()))
    almost_all_equal(ga1, ga2, 0.001)
    almost_all_equal(ga1, est, 0.001)


@pytest.mark.parametrize("through_arrays", [False, True])
def test_hessian_and_more_on_device(ctx, through_arrays):
    """tests/graph.rs:28-69 / :107-153 on GraphN over device values: exact f32, absent stays absent."""
    h = GOLDEN["hessian_f32"]
    g = gb.GraphN(ctx)
    x, y = g.variable(ctx.scalar(h["x"])), g.variable(ctx.scalar(h["y"]))
    if through_arrays:
        xa, ya = g.constant_as(x, (1,)), g.constant_as(y, (1,))
        z = g.as_scalar(g.matmul_nn(g.matmul_nn(xa, ya), ya))
    else:
        z = g.mul(g.mul(x, y), y)
    dz = g.compute_gradients(z.gid(), g.constant(ctx.scalar(1.0)))
    dz_dx, dz_dy = dz.get(x.gid()), dz.get(y.gid())
    ddx = g.compute_gradients(dz_dx.gid(), g.constant(ctx.scalar(1.0)))
    ddy = g.compute_gradients(dz_dy.gid(), g.constant(ctx.scalar(1.0)))
    f32 = np.float32
    assert ddx.get(x.gid()) is None
    assert f32(ddx.get(y.gid()).data().item()) == f32(h["ddz_dxdy"])
    assert f32(ddy.get(x.gid()).data().item()) == f32(h["ddz_dydx"])
    assert f32(ddy.get(y.gid()).data().item()) == f32(h["ddz_dydy"])
    dddx = g.compute_gradients(ddx.get(y.gid()).gid(), g.constant(ctx.scalar(1.0)))
    assert dddx.get(x.gid()) is None
    assert f32(dddx.get(y.gid()).data().item()) == f32(h["dddz_dxdydy"])


def test_array_ops_vs_finite_differences(ctx):
    """The reference's array tests (tests/{core,arith,analytic,array,matrix}.rs af modules): one op on a
    4x3 input, constant direction, L-inf < 1e-3 (2e-3 matmul) against central differences."""
    d43 = np.ones((4, 3), np.float32)
    cases = [
        ("exp", lambda g, v: g.exp(v), (4, 3), d43, 1e-3),
        ("tanh", lambda g, v: g.tanh(v), (4, 3), d43, 1e-3),
        ("neg", lambda g, v: g.neg(v), (4, 3), d43, 1e-3),
        ("powc", lambda g, v: g.powc(v, 4), (4, 3), d43, 1e-3),
        ("flat", lambda g, v: g.flat(v), (4, 3), np.ones((12,), np.float32), 1e-3),
        ("moddims", lambda g, v: g.moddims(v, (2, 2, 3)), (4, 3), np.ones((2, 2, 3), np.float32), 1e-3),
        ("sum_as", lambda g, v: g.sum_as(v, (1, 3)), (4, 3), np.ones((1, 3), np.float32), 1e-3),
        ("tile_as", lambda g, v: g.tile_as(v, (4, 3)), (1, 3), d43, 1e-3),
        ("transpose", lambda g, v: g.transpose(v), (4, 3), np.ones((3, 4), np.float32), 1e-3),
        ("max_as", lambda g, v: g.max_as(v, (1, 3)), (4, 3), np.ones((1, 3), np.float32), 1e-3),
    ]
    for name, f, shape, direction, prec in cases:
        x = rand(shape, 31)
        g = gb.Graph1(ctx)
        a = g.variable(x)
        b = f(g, a)
        grads = g.evaluate_gradients_once(b.gid(), direction)
        est = estimate_gradient(ctx, x, direction, 0.001, f)
        almost_all_equal(to_np(grads.get(a.gid())), est, prec)
    # tests/matrix.rs:25-49
    xa, xb = rand((4, 3), 32), rand((3, 5), 33)
    g = gb.Graph1(ctx)
    a, b = g.variable(xa), g.variable(xb)
    c = g.matmul_nn(a, b)
    direction = np.ones((4, 5), np.float32)
    grads = g.evaluate_gradients_once(c.gid(), direction)
    est_a = estimate_gradient(ctx, xa, direction, 0.001, lambda ev, v: ev.matmul_nn(v, ev.constant(xb)))
    est_b = estimate_gradient(ctx, xb, direction, 0.001, lambda ev, v: ev.matmul_nn(ev.constant(xa), v))
    almost_all_equal(to_np(grads.get(a.gid())), est_a, 0.002)
    almost_all_equal(to_np(grads.get(b.gid())), est_b, 0.002)


def test_training_converges_like_tests_net_rs(ctx):
    """tests/net.rs:91-114: train a 3x3 linear map with apply_gradient_step(-0.01) on one sample until
    loss < 1e-6, then the map sends `a` to the identity within 1e-2. The step is the reference's:
    fresh Graph1 per example, forward, evaluate_gradients_once(loss, 1), weights += lambda * grad."""
    a = np.array([1.0, 2.0, 1.0, 1.0, 0.0, 1.0, 0.0, -2.0, -1.0], np.float32).reshape((3, 3), order="F")
    ident = np.eye(3, dtype=np.float32)
    rng = np.random.default_rng(5)
    weights = ctx.array(rng.standard_normal((3, 3)).astype(np.float32))
    da, di = ctx.array(a), ctx.array(ident)
    loss = None
    for step in range(20000):
        g = gb.Graph1(ctx)
        inp = g.constant(da)
        w = g.variable(weights)
        out = g.matmul_nn(inp, w)
        delta = g.sub(g.constant(di), out)  # SquareLoss: net_ext.rs:110-118
        l = g.norm2(delta)
        store = g.evaluate_gradients_once(l.gid(), 1.0)
        weights.add_assign(store.get(w.gid()).scale(-0.01))  # WeightOps: net.rs:161-180
        if step % 50 == 0:
            loss = l.data().item()
            assert np.isfinite(loss)
            if loss < 1e-6:
                break
    assert loss is not None and loss < 1e-6
    ev = gb.Eval(ctx)
    i2 = to_np(ev.matmul_nn(ev.constant(da), ev.constant(weights)))
    almost_all_equal(i2, ident, 0.01)
    # check() returns the evaluated dims (tests/net.rs:150-151)
    assert gb.Check().matmul_nn(da.dims(), weights.dims()) == (3, 3, 1, 1)


def test_user_defined_operator(ctx):
    """tests/extension.rs:61-107: a `square` op defined outside the library through make_node, with two
    add_gradient calls; gradient of 3^2 is 6 — for Graph1 and, recorded, for GraphN."""
    for cls in (gb.Graph1, gb.GraphN):
        g = cls(ctx)
        a = g.variable(ctx.scalar(3.0))
        ev = gb.Eval(ctx)
        data = ev.mul(gb.Value(a.data()), gb.Value(a.data())).data()

        def vjp(ga, add_gradient, gradient, a=a):
            c = gb.Value(a.data(), a.id()) if cls is gb.GraphN else gb.Value(a.data())
            add_gradient(a.gid(), ga.mul(gradient, c))
            add_gradient(a.gid(), ga.mul(c, gradient))

        b = g.make_node(data, [a], vjp)
        assert b.data().item() == 9.0
        if cls is gb.Graph1:
            store = g.evaluate_gradients_once(b.gid(), 1.0)
            assert store.get(a.gid()).item() == 6.0
        else:
            store = g.compute_gradients(b.gid(), g.constant(ctx.scalar(1.0)))
            assert store.get(a.gid()).data().item() == 6.0
            assert store.get(a.gid()).id() is not None  # a recorded value: differentiable again
