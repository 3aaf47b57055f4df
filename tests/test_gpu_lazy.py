"""GPU: the sweep compiler — deferred elementwise regions compiled through NVRTC (csrc/batched.cu, jit.cu).

A deferred region computes, per element, exactly the rounded values of the eager kernels (same op order,
--fmad=false, same libdevice), so everything here compares with == against the eager one-kernel-per-primitive
sweep (fusion off), then with the stated tolerances against the CPU oracle. `ctx.set_lazy(True, 2)` forces
the deferred path on the small shapes the oracle finishes quickly; the default threshold is 32768 elements.
"""
import numpy as np
import pytest

import gad_b200 as gb
from gad_b200 import workloads as wl
from helpers import rand, to_np
from test_gpu_ops import CHAINS, LIBM_UNARY, assert_close, direction_seed, ones_like_seed, run_both

pytestmark = pytest.mark.gpu


@pytest.fixture
def lazy_ctx(ctx):
    ctx.set_lazy(True, 2)
    yield ctx
    ctx.set_lazy(True, 32768)
    ctx.set_fusion(True)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(4, 3), (4099, 3)], ids=str)
@pytest.mark.parametrize("name", sorted(CHAINS))
def test_deferred_sweep_is_bit_identical_to_the_eager_op_sequence(lazy_ctx, oracle, name, shape, dt):
    if name == "reduce_tile" and shape != (4, 3):
        pytest.skip("chain is written for [4,3]")
    a, b = rand(shape, 21, 0.5, 2.0, dt), rand(shape, 22, 0.5, 2.0, dt)
    lazy = run_both(lazy_ctx, oracle, CHAINS[name], [a, b], direction_seed, fuse=True)
    plain = run_both(lazy_ctx, oracle, CHAINS[name], [a, b], direction_seed, fuse=False)
    assert_close(lazy["fwd"], plain["fwd"], 0)
    for x, y in zip(lazy["grads"], plain["grads"]):
        assert_close(x, y, 0)
    for x, y in zip(lazy["grads"], lazy["ograds"]):
        assert np.allclose(x, y, rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("op", LIBM_UNARY + ["neg", "reciprocal", "sqrt", "abs", "relu"])
def test_deferred_unary_forward_and_vjp(lazy_ctx, oracle, op, dt):
    x = rand((1027,), 11, 0.25, 2.0, dt)
    if op in ("abs", "relu", "neg"):
        x = x - 1.0
    lazy = run_both(lazy_ctx, oracle, lambda g, v: getattr(g, op)(v[0]), [x], direction_seed, fuse=True)
    plain = run_both(lazy_ctx, oracle, lambda g, v: getattr(g, op)(v[0]), [x], direction_seed, fuse=False)
    assert_close(lazy["fwd"], plain["fwd"], 0)
    assert_close(lazy["grads"][0], plain["grads"][0], 0)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("shape,rd", [((1_000_003,), (1,)), ((4 * 1184 * 8192 // 16 + 4,), (1,)), ((40_000,), (1,)), ((4096, 37), (1, 37)),
                                      ((4099, 5), (1, 5)), ((128, 64, 3), (1, 1, 3)), ((128, 64, 3), (1, 1, 1)), ((6, 1), (1, 1))], ids=str)
@pytest.mark.parametrize("op", ["sum_as", "max_as"])
def test_reduction_of_a_deferred_value_is_fused_and_bit_identical(lazy_ctx, op, shape, rd, dt):
    """sum_as / max_as over the leading axes of a value that is still a deferred chain: one kernel computes the chain and
    reduces it in reduce.cu's own order (same cut into chunks, same order inside a chunk) — the bits are those of
    materialise-then-reduce, the value itself is never written (no launch of the elementwise region)."""
    x, y = rand(shape, 5, -1.0, 1.0, dt), rand(shape, 6, 0.5, 2.0, dt)
    dx, dy = lazy_ctx.array(x), lazy_ctx.array(y)

    def chain(ev):
        a, b = ev.constant(dx), ev.constant(dy)
        return getattr(ev, op)(ev.add(ev.mul(ev.exp(a), ev.log(b)), a), rd)

    ev = gb.Eval(lazy_ctx)
    chain(ev).data().numpy()  # compile
    l0 = lazy_ctx.launches()
    fused = chain(ev).data()
    launched = lazy_ctx.launches() - l0
    lazy_ctx.set_fusion(False)
    plain = chain(gb.Eval(lazy_ctx)).data().numpy()
    lazy_ctx.set_fusion(True)
    got = fused.numpy()
    assert got.dtype == dt and got.shape == plain.shape
    assert np.array_equal(got, plain), np.max(np.abs(got - plain))
    assert launched <= 2, launched  # the reducing kernel (+ the second stage when a segment was cut)
    ref = (np.exp(x.astype(np.float64)) * np.log(y.astype(np.float64)) + x).reshape(shape + (1,) * (4 - len(shape)), order="F")
    axes = tuple(i for i in range(len(rd)) if rd[i] != shape[i])
    want = ref.sum(axis=axes, keepdims=True) if op == "sum_as" else ref.max(axis=axes, keepdims=True)
    scale = np.abs(ref).sum(axis=axes, keepdims=True) if op == "sum_as" else np.abs(want)
    assert np.all(np.abs(got - want) <= 64 * np.finfo(dt).eps * scale + np.finfo(dt).tiny)


def test_fused_reduction_leaves_the_value_usable(lazy_ctx):
    """The reduced value stays a deferred chain: read afterwards (or consumed by a later op) it is computed then."""
    x = rand((70_001,), 8, -1.0, 1.0)
    ev = gb.Eval(lazy_ctx)
    d = ev.mul(ev.exp(ev.constant(lazy_ctx.array(x))), ev.constant(lazy_ctx.array(x)))
    s = ev.sum_as(d, (1,)).data().numpy()
    later = ev.addc(d, 1).data().numpy().reshape(-1)
    want = (np.exp(x) * x).astype(np.float32)
    assert np.allclose(later, want + 1, rtol=1e-6, atol=1e-6)
    assert np.array_equal(d.data().numpy().reshape(-1) + np.float32(1), later)
    assert abs(float(s.reshape(-1)[0]) - float(want.astype(np.float64).sum())) <= 64 * np.finfo(np.float32).eps * np.abs(want).sum()


def test_one_kernel_per_region(lazy_ctx):
    """C3: the forward is one compiled kernel that reduces what it computes, the whole reverse sweep is one compiled kernel that
    reads x, y and writes both gradients (the 4-pass algorithmic traffic)."""
    n = 1_000_003  # vector body + scalar tail
    x, y = rand((n,), 3, -1, 1), rand((n,), 4, 0.5, 2.0)
    dx, dy = lazy_ctx.array(x), lazy_ctx.array(y)
    wl.c3_step(lazy_ctx, dx, dy)  # compile
    lazy_ctx.synchronize()
    g = gb.Graph1(lazy_ctx)
    l0 = lazy_ctx.launches()
    vx, vy, z = wl.c3_forward(g, dx, dy)
    l1 = lazy_ctx.launches()
    store = g.evaluate_gradients_once(z.gid(), 1.0)
    l2 = lazy_ctx.launches()
    assert l1 - l0 <= 2, l1 - l0      # elementwise chain and reduction in ONE kernel + the second stage over its partials
    assert l2 - l1 == 2, l2 - l1      # the seed scalar's fill + ONE kernel for the sweep
    gx, gy = store.get(vx.gid()).numpy().reshape(-1), store.get(vy.gid()).numpy().reshape(-1)
    lazy_ctx.set_fusion(False)
    _, ex, ey = wl.c3_step(lazy_ctx, dx, dy)
    lazy_ctx.set_fusion(True)
    assert np.array_equal(gx, ex.numpy().reshape(-1)) and np.array_equal(gy, ey.numpy().reshape(-1))
    # intermediate gradients stay deferred until somebody reads them, and then are the eager values
    eps = np.finfo(np.float32).eps
    want_gy = np.exp(x.astype(np.float64)) / y.astype(np.float64)
    assert np.all(np.abs(gy - want_gy) <= 8 * eps * np.abs(want_gy))


@pytest.mark.parametrize("B,D", [(8, 4), (64, 32), (260, 24)])
def test_mlp_with_deferred_bias_tiles_matches_eager(lazy_ctx, B, D):
    """The bias `tile_as` stays a column-tile view that the compiled kernel indexes; tanh forward and VJP,
    the loss chain and the accumulations are deferred; GEMMs and reductions force their operands."""
    rng = np.random.default_rng(5)
    X = rng.standard_normal((B, D)).astype(np.float32)
    Ws = [(rng.standard_normal((D, D)) / np.sqrt(D)).astype(np.float32) for _ in range(3)]
    bs = [rng.standard_normal((1, D)).astype(np.float32) * 0.1 for _ in range(3)]
    T = rng.standard_normal((B, D)).astype(np.float32)

    def run():
        loss, gW, gb_ = wl.mlp_step(lazy_ctx, lazy_ctx.array(X), [lazy_ctx.array(w) for w in Ws], [lazy_ctx.array(b) for b in bs], lazy_ctx.array(T))
        return loss.item(), [g.numpy() for g in gW], [g.numpy() for g in gb_]

    lazy = run()
    lazy_ctx.set_fusion(False)
    plain = run()
    lazy_ctx.set_fusion(True)
    assert lazy[0] == plain[0]
    for a, b in zip(lazy[1] + lazy[2], plain[1] + plain[2]):
        assert np.array_equal(a, b)


def test_row_tile_and_uniform_operands(lazy_ctx, oracle):
    """[B,1] tiled over [B,D] (row tile), a scalar broadcast, and a dense operand in one region."""
    a, r = rand((37, 5), 1, 0.5, 2.0), rand((37, 1), 2, 0.5, 2.0)
    build = lambda g, v: g.mul(g.add(v[0], g.tile_as(v[1], (37, 5))), g.constant_as(g.as_scalar(g.sum_as(v[1], (1,))), (37, 5)))
    lazy = run_both(lazy_ctx, oracle, build, [a, r], direction_seed, fuse=True)
    plain = run_both(lazy_ctx, oracle, build, [a, r], direction_seed, fuse=False)
    assert_close(lazy["fwd"], plain["fwd"], 0)
    for x, y in zip(lazy["grads"], plain["grads"]):
        assert_close(x, y, 0)
    for x, y in zip(lazy["grads"], lazy["ograds"]):
        assert np.allclose(x, y, rtol=1e-5, atol=1e-5)


def test_in_place_write_forces_pending_readers_first(lazy_ctx):
    """WeightOps::add_assign (net.rs:171-175) overwrites a buffer a deferred value still wants to read: the value
    is computed from the OLD contents first, as the eager order of operations would have it."""
    w = rand((1000,), 1, -1, 1)
    d = rand((1000,), 2, -1, 1)
    ev = gb.Eval(lazy_ctx)
    dw, dd = lazy_ctx.array(w), lazy_ctx.array(d)
    y = ev.exp(ev.mulc(ev.constant(dw), 2.0))  # deferred, reads dw
    dw.add_assign(dd)
    got = to_np(y).reshape(-1)
    lazy_ctx.set_lazy(False, 2)
    want = to_np(ev.exp(ev.mulc(ev.constant(lazy_ctx.array(w)), 2.0))).reshape(-1)
    assert np.array_equal(got, want)
    assert np.array_equal(dw.numpy().reshape(-1), w + d)


def test_graphn_records_through_deferred_regions(lazy_ctx, oracle):
    """GraphN's gradient algebra is the graph: its recorded backward ops go through the same deferred Eval."""
    x, y = rand((50,), 1, 0.5, 1.5), rand((50,), 2, 0.5, 1.5)

    def hess(ctx_fuse):
        lazy_ctx.set_fusion(ctx_fuse)
        g = gb.GraphN(lazy_ctx)
        vx, vy = g.variable(x), g.variable(y)
        z = g.mul(vx, g.mul(vy, vy))  # z = x y^2 (tests/graph.rs:28-69 on arrays)
        first = g.compute_gradients(z.gid(), g.constant(np.ones(50, np.float32)))
        dzdy = first.get(vy.gid())
        second = g.compute_gradients(dzdy.gid(), g.constant(np.ones(50, np.float32)))
        out = to_np(second.get(vy.gid()).data()).reshape(-1), to_np(second.get(vx.gid()).data()).reshape(-1)
        lazy_ctx.set_fusion(True)
        return out

    a, b = hess(True), hess(False)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.allclose(a[0], 2 * x, rtol=1e-6) and np.allclose(a[1], 2 * y, rtol=1e-6)


@pytest.mark.parametrize("n_rows,n_cols", [((1 << 22) + 3, 1), (4099, 1031), (8192, 1024)])
def test_large_regions_take_the_ring_fed_kernel_and_keep_their_bits(ctx, n_rows, n_cols):
    """Regions of >= 2^22 elements stream their dense inputs through a shared-memory ring (cp.async.bulk producer warp):
    same values as the eager one-kernel-per-op path, for a length that is not a multiple of the vector width, with a
    column-tiled operand, several dense inputs and two outputs."""
    rng = np.random.default_rng(21)
    shape = (n_rows, n_cols) if n_cols > 1 else (n_rows,)
    x = rng.uniform(-1, 1, shape).astype(np.float32)
    y = rng.uniform(0.5, 2, shape).astype(np.float32)
    w = rng.uniform(0.5, 1.5, shape).astype(np.float32)
    b = rng.uniform(-1, 1, (1, n_cols)).astype(np.float32)

    def run(lazy):
        ctx.set_lazy(lazy)
        try:
            ev = gb.Eval(ctx)
            vx, vy, vw = ev.constant(x), ev.constant(y), ev.constant(w)
            t = ev.tanh(ev.add(vx, ev.tile_as(ev.constant(b), shape))) if n_cols > 1 else ev.tanh(vx)
            u = ev.mul(ev.exp(t), ev.log(vy))
            v = ev.div(ev.add(u, vw), vy)
            return to_np(u), to_np(v)
        finally:
            ctx.set_lazy(True)

    a, e = run(True), run(False)
    assert np.array_equal(a[0], e[0]) and np.array_equal(a[1], e[1])
