"""GPU: the BASELINE.json configurations against the oracle at sizes the oracle finishes in seconds,
and through size-independent properties at full size."""
import numpy as np
import pytest

import gad_b200 as gb
import programs
from gad_b200 import workloads as wl
from helpers import rand, to_np, ulp_diff

pytestmark = pytest.mark.gpu


# ---- C1 / C2: scalar tapes, one per lane ------------------------------------------------------------
@pytest.mark.parametrize("n,lanes", [(2, 1), (8, 33), (64, 1000), (64, 4097), (64, 256 * 148 * 5 + 8), (40, 4)])
def test_batched_rosenbrock_matches_scalar_tape_bit_for_bit(ctx, oracle, n, lanes):
    """Rosenbrock uses only +, -, *, neg and constants: every op is an exactly rounded IEEE operation,
    the op order is the tape's, no FMA contraction on either side => f64 results are bit-identical to
    the reference-order scalar tape run lane by lane on the CPU."""
    rng = np.random.default_rng(2)
    x = -1.2 + 2.4 * rng.random((n, lanes))
    cols = [ctx.array(x[i]) for i in range(n)]
    g, f, grads, store = wl.rosenbrock_batched_gradients(ctx, cols)
    of, ograd, _ = oracle.rosenbrock_batch(x, nthreads=4)
    got = np.stack([to_np(c)[:, 0, 0, 0] for c in grads])
    assert got.dtype == np.float64
    assert np.array_equal(got, ograd), f"max ulp {ulp_diff(got, ograd)}"
    assert np.allclose(got, programs.rosenbrock_grad_closed_form(x), rtol=1e-11, atol=1e-9)
    fv = to_np(g.force(f))[:, 0, 0, 0]
    assert np.array_equal(fv, of)
    assert g.num_nodes() == n + 8 * (n - 1) + 1  # same tape as the scalar Graph1 (569 for N=64)
    stats = store.program_stats()
    assert stats["instructions"] > 0 and stats["slots"] > 0
    # every tracked node has a store entry, as after a scalar sweep
    assert len(store.ids()) == g.num_nodes()


def test_batched_rerun_at_new_points_equals_a_fresh_tape(ctx, oracle):
    """gadcu_batched_rerun: the compiled lane program at new columns == recording the tapes again == the scalar tape."""
    n, lanes = 16, 777
    rng = np.random.default_rng(7)
    x0, x1 = -1.2 + 2.4 * rng.random((n, lanes)), -1.2 + 2.4 * rng.random((n, lanes))
    g, f, grads, store = wl.rosenbrock_batched_gradients(ctx, [ctx.array(x0[i]) for i in range(n)])
    launches = ctx.launches()
    again = store.rerun([ctx.array(x1[i]) for i in range(n)])
    assert ctx.launches() - launches == 1  # one kernel, nothing re-recorded
    got = np.stack([to_np(c)[:, 0, 0, 0] for c in again])
    _, _, fresh, _ = wl.rosenbrock_batched_gradients(ctx, [ctx.array(x1[i]) for i in range(n)])
    assert np.array_equal(got, np.stack([to_np(c)[:, 0, 0, 0] for c in fresh]))
    _, ograd, _ = oracle.rosenbrock_batch(x1, nthreads=4)
    assert np.array_equal(got, ograd)
    # the first sweep's results are untouched
    _, ograd0, _ = oracle.rosenbrock_batch(x0, nthreads=4)
    assert np.array_equal(np.stack([to_np(c)[:, 0, 0, 0] for c in grads]), ograd0)
    with pytest.raises(gb.LengthsError):
        store.rerun([ctx.array(x1[0])])


def test_batched_sweep_over_a_known_structure_is_replayed_not_walked(ctx, oracle):
    """The sweep memo (batched.cu): a second tape of the same structure gets the first one's backward instructions and
    store layout without walking the nodes — same bits, same store ids, same consumption of the tape; a tape that
    differs in one constant, in the root, or in the kind of sweep is walked."""
    n, lanes = 12, 515
    rng = np.random.default_rng(11)
    xs = [-1.2 + 2.4 * rng.random((n, lanes)) for _ in range(3)]

    def record(x, scale=100, once=True, root_last=True):
        g = gb.BatchedGraph1(ctx, lanes, gb.F64)
        vs = [g.variable(ctx.array(x[i])) for i in range(n)]
        terms = []
        for i in range(n - 1):
            d = g.sub(vs[i + 1], g.mul(vs[i], vs[i]))
            e = g.addc(g.neg(vs[i]), 1)
            terms.append(g.add(g.mulc(g.mul(d, d), scale), g.mul(e, e)))
        f = g.add_all(terms if root_last else terms[:-1])
        store = (g.evaluate_gradients_once if once else g.evaluate_gradients)(f.gid(), 1.0)
        inner.append(terms[:2])
        return g, f, vs, store

    def grads(vs, store):
        got = [store.get(v.gid()) for v in vs]  # (None: a variable the root does not depend on)
        return np.stack([to_np(c)[:, 0, 0, 0] if c is not None else np.zeros(lanes) for c in got])

    inner = []  # (two intermediate nodes of every recorded tape)
    s0 = ctx.sweep_memo_stats()
    g1, f1, v1, st1 = record(xs[0])
    s1 = ctx.sweep_memo_stats()
    g2, f2, v2, st2 = record(xs[1])
    s2 = ctx.sweep_memo_stats()
    assert s1["walked"] - s0["walked"] + s1["replayed"] - s0["replayed"] == 1
    assert (s2["replayed"] - s1["replayed"], s2["walked"] - s1["walked"]) == (1, 0)
    for x, vs, st in ((xs[0], v1, st1), (xs[1], v2, st2)):
        _, ograd, _ = oracle.rosenbrock_batch(x, nthreads=2)
        assert np.array_equal(grads(vs, st), ograd)
    assert st1.ids() == st2.ids()
    assert len(st2.ids()) == g2.num_nodes()
    # entries nobody asked the sweep for (the root's own gradient, an intermediate node's) are there when read, in the
    # replayed store as in the walked one, and reading them does not change what the store lists
    for g, st, f, terms in ((g1, st1, f1, inner[0]), (g2, st2, f2, inner[1])):
        column = lambda entry: to_np(g.force(gb.Value(entry)))[:, 0, 0, 0]  # (a lane value is read through gadcu_batched_force)
        assert np.array_equal(column(st.get(f.gid())), np.ones(lanes))
        for t in terms:
            assert np.array_equal(column(st.get(t.gid())), np.ones(lanes))  # d(sum of terms)/d(term)
        with pytest.raises(gb.InvalidError):
            st.get(terms[0].gid()).numpy()  # its dims are one tape's scalar: the column does not fit a host buffer sized by them
        assert st.get(gb.GradientId(f.gid().arena, g1.num_nodes() + 5)) is None
    assert st1.ids() == st2.ids() and len(st2.ids()) == g2.num_nodes()
    assert st2.program_stats() == st1.program_stats()
    # the replayed sweep consumed the tape like the walked one: a second sweep finds the root without its update
    for g, f in ((g1, f1), (g2, f2)):
        again = g.evaluate_gradients_once(f.gid(), 1.0)
        assert len(again.ids()) == 1
    # the re-run of a replayed sweep's program
    re = st2.rerun([ctx.array(xs[2][i]) for i in range(n)])
    _, ograd2, _ = oracle.rosenbrock_batch(xs[2], nthreads=2)
    assert np.array_equal(np.stack([to_np(c)[:, 0, 0, 0] for c in re]), ograd2)
    # other structures are walked: another constant, another root, a non-consuming sweep (which can be swept twice)
    for kw in ({"scale": 99}, {"root_last": False}, {"once": False}):
        before = ctx.sweep_memo_stats()
        g, f, vs, st = record(xs[0], **kw)
        after = ctx.sweep_memo_stats()
        assert after["walked"] - before["walked"] == 1, kw
        g_, f_, vs_, st_ = record(xs[0], **kw)
        assert ctx.sweep_memo_stats()["replayed"] - after["replayed"] == 1, kw
        assert np.array_equal(grads(vs, st), grads(vs_, st_))
        if kw.get("scale") == 99:
            assert not np.array_equal(grads(vs, st), grads(v1, st1))
        if kw.get("once") is False:
            twice = g_.evaluate_gradients(f_.gid(), 1.0)
            assert np.array_equal(grads(vs_, twice), grads(vs_, st_))


def test_wide_batched_tape_beyond_4k_of_kernel_parameters(ctx, oracle):
    """Rosenbrock-320: 320 input and 320 output columns (5 KB of pointers) + 640 constants in one lane program."""
    n, lanes = 320, 64
    x = -1.2 + 2.4 * np.random.default_rng(11).random((n, lanes))
    g, f, grads, store = wl.rosenbrock_batched_gradients(ctx, [ctx.array(x[i]) for i in range(n)])
    got = np.stack([to_np(c)[:, 0, 0, 0] for c in grads])
    _, ograd, _ = oracle.rosenbrock_batch(x, nthreads=4)
    assert np.array_equal(got, ograd)


def test_batched_tape_golden_point_and_f32(ctx):
    cols = [ctx.array(np.array([-1.2, 1.0])), ctx.array(np.array([1.0, 1.0]))]
    g, f, grads, _ = wl.rosenbrock_batched_gradients(ctx, cols)
    got = np.stack([to_np(c)[:, 0, 0, 0] for c in grads])
    assert np.allclose(got[:, 0], [-215.6, -88.0], atol=1e-10)
    assert np.all(got[:, 1] == 0.0)
    # analytic ops through the lane interpreter (f32), against the oracle's scalar Eval
    x = rand((257,), 4, 0.5, 2.0)
    for name in ["exp", "log", "log1p", "sin", "cos", "tanh", "sigmoid", "reciprocal", "sqrt"]:
        bg = gb.BatchedGraph1(ctx, 257, gb.F32)
        v = bg.variable(ctx.array(x))
        out = getattr(bg, name)(bg.mulc(v, 2))
        store = bg.evaluate_gradients_once(out.gid(), 1.7)
        got = to_np(store.get(v.gid()))[:, 0, 0, 0]
        x64 = x.astype(np.float64) * 2
        d = {"exp": np.exp(x64), "log": 1 / x64, "log1p": 1 / (1 + x64), "sin": np.cos(x64), "cos": -np.sin(x64),
             "tanh": 1 - np.tanh(x64) ** 2, "sigmoid": np.exp(-x64) / (1 + np.exp(-x64)) ** 2, "reciprocal": -1 / x64 ** 2,
             "sqrt": 0.5 / np.sqrt(x64)}[name] * 2 * 1.7
        assert np.allclose(got, d, rtol=2e-5, atol=1e-6), name


def test_staged_lane_program_f32_f64_partial_tiles(ctx):
    """Long tapes stream their input columns through a shared-memory ring fed by cp.async.bulk (batched.cu, JitPlan::ring):
    f64 and f32, whole and partial last tiles, against the tape's reverse sweep restated in numpy (the f64 bit-for-bit
    comparison with the scalar oracle tape is test_batched_rosenbrock_matches_scalar_tape_bit_for_bit, whose 1000-,
    4-lane and 189 448-lane cases run through this kernel too); a re-run gives the same bits."""
    for dtype, npdt, lanes in [(gb.F64, np.float64, 2000), (gb.F32, np.float32, 1004), (gb.F32, np.float32, 128 * 2 * 3)]:
        n = 48
        x = (-1.2 + 2.4 * np.random.default_rng(lanes).random((n, lanes))).astype(npdt)
        cols = [ctx.array(x[i]) for i in range(n)]
        _, f, grads, store = wl.rosenbrock_batched_gradients(ctx, cols, dtype)
        got = np.stack([to_np(c)[:, 0, 0, 0] for c in grads])
        assert got.dtype == npdt
        # the same tape, one lane at a time through numpy in the tape's operation order
        gx = np.zeros_like(x)
        for i in range(n - 1):
            sq = x[i] * x[i]; d = x[i + 1] - sq; e = -x[i] + npdt(1)
            # forward: t1 = 100 d^2, e2 = e^2; reverse (seed 1): see SURVEY App. D
            gd = (npdt(1) * npdt(100)) * d + (npdt(1) * npdt(100)) * d   # d2 = d*d: both operand contributions
            ge = e + e
            gx[i + 1] += gd
            gx[i] += -(gd * x[i] + gd * x[i]) - ge
        assert np.allclose(got, gx, rtol=2e-4 if npdt == np.float32 else 1e-12, atol=1e-3 if npdt == np.float32 else 1e-9)
        again = store.rerun(cols)
        assert np.array_equal(np.stack([to_np(c)[:, 0, 0, 0] for c in again]), got)


def test_batched_select_pow_div_vs_scalar_oracle(ctx, oracle):
    lanes = 65
    a, b = rand((lanes,), 1, 0.5, 2.0, np.float64), rand((lanes,), 2, 0.5, 2.0, np.float64)
    bg = gb.BatchedGraph1(ctx, lanes, gb.F64)
    va, vb = bg.variable(ctx.array(a)), bg.variable(ctx.array(b))
    out = bg.add(bg.max(va, vb), bg.mul(bg.div(va, vb), bg.powc(vb, 3)))
    store = bg.evaluate_gradients_once(out.gid(), 1.0)
    ga, gb_ = to_np(store.get(va.gid()))[:, 0, 0, 0], to_np(store.get(vb.gid()))[:, 0, 0, 0]
    for lane in range(lanes):
        og = oracle.OGraph1()
        oa, ob = og.variable(np.float64(a[lane])), og.variable(np.float64(b[lane]))
        oout = og.add(og.max(oa, ob), og.mul(og.div(oa, ob), og.powc(ob, 3)))
        ost = og.evaluate_gradients_once(oout.gid(), np.float64(1.0))
        assert ulp_diff(np.float64(ga[lane]), np.float64(ost.get(oa.gid(), like=oa).data())) <= 2
        assert ulp_diff(np.float64(gb_[lane]), np.float64(ost.get(ob.gid(), like=ob).data())) <= 2


# ---- C3: elementwise tape ------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 1000, 2 ** 20 + 3])
def test_c3_matches_oracle(ctx, oracle, n):
    x, y = rand((n,), 3, -1.0, 1.0), rand((n,), 4, 0.5, 2.0)
    z, gx, gy = wl.c3_step(ctx, ctx.array(x), ctx.array(y))
    oz, ogx, ogy, _, _ = oracle.c3(x, y)
    # d/dx = exp(x)*log(y) + 1 cancels near exp(x)*log(y) = -1: bound in terms of the summands
    # (4 ulp of libm slack on the product + the rounding of the sum); d/dy = exp(x)/y: 4 ulp
    eps = np.finfo(np.float32).eps
    prod = np.abs(np.exp(x.astype(np.float64)) * np.log(y.astype(np.float64)))
    assert np.all(np.abs(to_np(gx).reshape(-1).astype(np.float64) - ogx) <= 6 * eps * (prod + 1.0))
    assert ulp_diff(to_np(gy).reshape(-1), ogy) <= 4
    terms = np.exp(x.astype(np.float64)) * np.log(y.astype(np.float64)) + x
    assert abs(z.item() - terms.sum()) <= 8 * np.finfo(np.float32).eps * np.abs(terms).sum()
    assert abs(oz - terms.sum()) <= 64 * n * np.finfo(np.float32).eps * max(1.0, np.abs(terms).max())


def test_c3_full_size_properties(ctx):
    """n = 64 Mi (BASELINE config 3): too big for the scalar oracle in seconds, so check properties:
    closed form on a strided sample, seed linearity (x2 is exact in binary), fused == unfused."""
    n = 64 * 1024 * 1024
    x = ctx.random((n,), 3, a=-1.0, b=1.0)
    y = ctx.random((n,), 4, a=0.5, b=2.0)
    z, gx, gy = wl.c3_step(ctx, x, y)
    hx, hy = x.numpy().reshape(-1)[:: 4099], y.numpy().reshape(-1)[:: 4099]
    hgx, hgy = gx.numpy().reshape(-1), gy.numpy().reshape(-1)
    want_gx = (np.exp(hx.astype(np.float64)) * np.log(hy.astype(np.float64)) + 1.0).astype(np.float32)
    want_gy = (np.exp(hx.astype(np.float64)) / hy.astype(np.float64)).astype(np.float32)
    eps = np.finfo(np.float32).eps
    prod = np.abs(np.exp(hx.astype(np.float64)) * np.log(hy.astype(np.float64)))
    assert np.all(np.abs(hgx[:: 4099].astype(np.float64) - want_gx) <= 6 * eps * (prod + 1.0))
    assert ulp_diff(hgy[:: 4099], want_gy) <= 4
    # seed linearity: sweep with seed 2.0 == 2 * sweep with seed 1.0, exactly
    g = gb.Graph1(ctx)
    vx, vy, zz = wl.c3_forward(g, x, y)
    s2 = g.evaluate_gradients(zz.gid(), 2.0)
    assert np.array_equal(s2.get(vx.gid()).numpy().reshape(-1), 2.0 * hgx)
    del s2
    # the unfused (reference op sequence) sweep gives the same bits
    ctx.set_fusion(False)
    s1 = g.evaluate_gradients_once(zz.gid(), 1.0)
    ctx.set_fusion(True)
    assert np.array_equal(s1.get(vy.gid()).numpy().reshape(-1), hgy)
    assert np.isfinite(z.item())


# ---- C4: MLP -------------------------------------------------------------------------------------
def _mlp_inputs(B, D, L=3, seed=5):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((B, D)).astype(np.float32)
    Ws = [(rng.standard_normal((D, D)) / np.sqrt(D)).astype(np.float32) for _ in range(L)]
    bs = [np.full((1, D), 0.01, np.float32) for _ in range(L)]
    T = rng.standard_normal((B, D)).astype(np.float32)
    return X, Ws, bs, T


def _mlp_reference_f64(X, Ws, bs, T):
    h = X.astype(np.float64)
    acts = [h]
    for l, (W, b) in enumerate(zip(Ws, bs)):
        z = h @ W.astype(np.float64) + b.astype(np.float64)
        h = np.tanh(z) if l + 1 < len(Ws) else z
        acts.append(h)
    delta = T.astype(np.float64) - h
    loss = np.sum(delta * delta)
    g = -2.0 * delta
    gWs, gbs = [None] * len(Ws), [None] * len(Ws)
    for l in reversed(range(len(Ws))):
        if l + 1 < len(Ws):
            g = g * (1 - acts[l + 1] ** 2)
        gWs[l] = acts[l].T @ g
        gbs[l] = g.sum(axis=0, keepdims=True)
        g = g @ Ws[l].astype(np.float64).T
    return loss, gWs, gbs


@pytest.mark.parametrize("B,D", [(8, 4), (64, 32), (256, 128), (512, 256)])
def test_mlp_step_matches_oracle(ctx, oracle, B, D):
    X, Ws, bs, T = _mlp_inputs(B, D)
    loss, gW, gb_ = wl.mlp_step(ctx, ctx.array(X), [ctx.array(w) for w in Ws], [ctx.array(b) for b in bs], ctx.array(T))
    oloss, ogW, ogb, _ = oracle.mlp_step(X, Ws, bs, T, nthreads=8)
    rloss, rgW, rgb = _mlp_reference_f64(X, Ws, bs, T)
    # both the CUDA path and the oracle sit within f32 rounding of the f64 reference; the CUDA path may
    # not be further from it than a few times the oracle's own error
    assert abs(loss.item() - rloss) <= 1e-5 * abs(rloss)
    assert abs(oloss - rloss) <= 1e-4 * abs(rloss)
    for l in range(3):
        a, o, r = to_np(gW[l])[:, :, 0, 0].astype(np.float64), ogW[l].astype(np.float64), rgW[l]
        scale = np.max(np.abs(r))
        assert np.max(np.abs(a - r)) <= 2e-5 * scale * np.sqrt(B), (l, np.max(np.abs(a - r)), scale)
        assert np.max(np.abs(a - r)) <= 8 * np.max(np.abs(o - r)) + 2e-6 * scale  # tcgen05 accumulates with truncation
        a, r = to_np(gb_[l])[:, :, 0, 0].astype(np.float64), rgb[l]
        assert np.max(np.abs(a - r)) <= 2e-5 * np.max(np.abs(r)) * np.sqrt(B)


def test_mlp_sharded_rows_sum_to_the_full_batch(ctx):
    """The data-parallel seam: gradients of row shards add up to the full-batch gradient."""
    B, D = 256, 64
    X, Ws, bs, T = _mlp_inputs(B, D)
    dW, db = [ctx.array(w) for w in Ws], [ctx.array(b) for b in bs]
    loss, gW, gb_ = wl.mlp_step(ctx, ctx.array(X), dW, db, ctx.array(T))
    full = to_np(gW[0])
    parts, lsum = None, 0.0
    for r in range(4):
        rows = slice(r * B // 4, (r + 1) * B // 4)
        l, gWr, _ = wl.mlp_step(ctx, ctx.array(X[rows]), dW, db, ctx.array(T[rows]))
        lsum += l.item()
        if parts is None:
            parts = gWr[0]
        else:
            parts.add_assign(gWr[0])
    assert np.allclose(to_np(parts), full, rtol=1e-4, atol=1e-5 * np.max(np.abs(full)))
    assert abs(lsum - loss.item()) <= 1e-5 * abs(loss.item())


def test_mlp_hvp_matches_finite_difference_of_gradients(ctx):
    """C5 (GraphN, reverse-over-reverse): H v ~ (grad(W + eps v) - grad(W - eps v)) / 2 eps."""
    B, D = 32, 16
    X, Ws, bs, T = _mlp_inputs(B, D)
    rng = np.random.default_rng(10)
    vs = [rng.standard_normal((D, D)).astype(np.float32) for _ in range(3)]
    dX, dT, db = ctx.array(X), ctx.array(T), [ctx.array(b) for b in bs]
    hv, loss, nodes = wl.mlp_hvp(ctx, dX, [ctx.array(w) for w in Ws], db, dT, [ctx.array(v) for v in vs])
    assert nodes > 19  # the recorded backward passes grew the tape
    eps = 1e-2
    _, gp, _ = _mlp_reference_f64(X, [w + eps * v for w, v in zip(Ws, vs)], bs, T)
    _, gm, _ = _mlp_reference_f64(X, [w - eps * v for w, v in zip(Ws, vs)], bs, T)
    for l in range(3):
        fd = (gp[l] - gm[l]) / (2 * eps)
        got = to_np(hv[l])[:, :, 0, 0]
        assert np.max(np.abs(got - fd)) <= 2e-2 * np.max(np.abs(fd)) + 1e-3


# ---- CUDA-graph capture --------------------------------------------------------------------------
def test_captured_step_replays_bit_identically(ctx):
    n = 100_003
    x, y = ctx.array(rand((n,), 3, -1, 1)), ctx.array(rand((n,), 4, 0.5, 2.0))
    z0, gx0, gy0 = wl.c3_step(ctx, x, y)  # eager (also warms the allocator)
    want_gx, want_gy, want_z = gx0.numpy(), gy0.numpy(), z0.item()
    with gb.Capture(ctx) as cap:
        z, gx, gy = wl.c3_step(ctx, x, y)
    for _ in range(3):
        cap.launch()
    ctx.synchronize()
    assert np.array_equal(gx.numpy(), want_gx) and np.array_equal(gy.numpy(), want_gy)
    assert z.item() == want_z
    # new inputs in the same buffers: the replay recomputes from them
    x.upload(np.ascontiguousarray(rand((n,), 30, -1, 1)))
    ctx.synchronize()
    cap.launch()
    z1, gx1, gy1 = wl.c3_step(ctx, x, y)
    assert np.array_equal(gx.numpy(), gx1.numpy())


def test_no_memory_growth_over_repeated_steps(ctx):
    """The caching pool, the lane programs (one per tape epoch) and the split cache release what a step used:
    after warm-up, 40 more mixed steps leave bytes-in-use and bytes-reserved unchanged."""
    B, D = 512, 256
    X, Ws, bs, T = _mlp_inputs(B, D)
    dX, dT = ctx.array(X), ctx.array(T)
    dW, db = [ctx.array(w) for w in Ws], [ctx.array(b) for b in bs]
    n = 1 << 17
    x, y = ctx.array(rand((n,), 3, -1, 1)), ctx.array(rand((n,), 4, 0.5, 2.0))
    cols = [ctx.array(-1.2 + 2.4 * np.random.default_rng(i).random(4096)) for i in range(8)]

    def one():
        wl.mlp_step(ctx, dX, dW, db, dT)[0].item()
        wl.c3_step(ctx, x, y)[0].item()
        g, f, grads, store = wl.rosenbrock_batched_gradients(ctx, cols)
        store.rerun(cols)

    for _ in range(4):
        one()
    ctx.synchronize()
    before = ctx.stats()
    for _ in range(40):
        one()
    ctx.synchronize()
    after = ctx.stats()
    assert after["bytes_in_use"] == before["bytes_in_use"] and after["bytes_reserved"] == before["bytes_reserved"], (before, after)


def test_one_tape_swept_from_several_threads(ctx):
    """gad's graphs are Send + Sync and evaluate_gradients(&self) does not modify the tape (lib.rs:69-72,
    graph.rs:263-274): four host threads sweep the same tape with different seeds at the same time."""
    import threading
    n = 100_003
    x, y = rand((n,), 3, -1, 1), rand((n,), 4, 0.5, 2.0)
    g = gb.Graph1(ctx)
    vx, vy, z = wl.c3_forward(g, ctx.array(x), ctx.array(y))
    want = g.evaluate_gradients(z.gid(), 1.0).get(vx.gid()).numpy()
    results, errors = {}, []

    def work(k):
        try:
            for _ in range(5):
                s = g.evaluate_gradients(z.gid(), float(2 ** k))
                results[k] = s.get(vx.gid()).numpy()
        except Exception as e:  # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k in range(4):
        assert np.array_equal(results[k], want * np.float32(2 ** k))


def test_captured_mlp_step_replays_bit_identically(ctx):
    """A whole C4 step — tcgen05 GEMMs (TMA maps, split kernels, split-K memsets), compiled elementwise regions,
    reductions — captured into one CUDA graph and replayed: same loss and gradients as the eager step, and the
    replay follows new contents of the input buffers."""
    B, D = 512, 256
    X, Ws, bs, T = _mlp_inputs(B, D)
    dX, dT = ctx.array(X), ctx.array(T)
    dW, db = [ctx.array(w) for w in Ws], [ctx.array(b) for b in bs]
    eager = wl.mlp_step(ctx, dX, dW, db, dT)
    want_loss, want_W = eager[0].item(), [g.numpy() for g in eager[1]]
    with gb.Capture(ctx) as cap:
        out = wl.mlp_step(ctx, dX, dW, db, dT)
    for _ in range(2):
        cap.launch()
    ctx.synchronize()
    assert out[0].item() == want_loss
    for a, b in zip(out[1], want_W):
        assert np.array_equal(a.numpy(), b)
    X2 = _mlp_inputs(B, D, seed=17)[0]
    dX.upload(np.ascontiguousarray(X2.reshape(-1, order="F")))   # upload takes the device (column-major) layout
    ctx.synchronize()
    cap.launch()
    fresh = wl.mlp_step(ctx, ctx.array(X2), dW, db, dT)
    assert out[0].item() == fresh[0].item()
    assert np.array_equal(out[1][2].numpy(), fresh[1][2].numpy())


def test_device_results_against_the_committed_oracle_fixtures(ctx):
    """tests/golden/oracle_array_fixtures.npz: the oracle's outputs for C3 (n = 257), C4 (B = 8, D = 4) and batched
    Rosenbrock (8 variables, 5 lanes) on seeded inputs, committed with the script that made them. Same bounds as the
    live-oracle tests above: batched f64 tapes bit for bit, f32 elementwise VJPs within a few ulp of the summands,
    GEMM-backed gradients within f32 rounding."""
    import os
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_array_fixtures.npz"))
    # C3
    x, y = fx["c3_x"], fx["c3_y"]
    z, gx, gy = wl.c3_step(ctx, ctx.array(x), ctx.array(y))
    eps = np.finfo(np.float32).eps
    prod = np.abs(np.exp(x.astype(np.float64)) * np.log(y.astype(np.float64)))
    assert np.all(np.abs(to_np(gx).reshape(-1).astype(np.float64) - fx["c3_gx"].astype(np.float64)) <= 12 * eps * (prod + 1.0))
    assert ulp_diff(to_np(gy).reshape(-1), fx["c3_gy"]) <= 4
    assert abs(z.item() - float(fx["c3_z"])) <= 64 * x.size * eps
    # C4
    X, Ws, bs, T = fx["mlp_X"], list(fx["mlp_W"]), list(fx["mlp_b"]), fx["mlp_T"]
    loss, gW, gb_ = wl.mlp_step(ctx, ctx.array(X), [ctx.array(w) for w in Ws], [ctx.array(b) for b in bs], ctx.array(T))
    assert abs(loss.item() - float(fx["mlp_loss"])) <= 1e-4 * abs(float(fx["mlp_loss"]))
    for l in range(3):
        a, o = to_np(gW[l])[:, :, 0, 0].astype(np.float64), fx["mlp_gW"][l].astype(np.float64)
        assert np.max(np.abs(a - o)) <= 4e-5 * np.max(np.abs(o)) * np.sqrt(X.shape[0]), l
        a, o = to_np(gb_[l])[:, :, 0, 0].astype(np.float64).reshape(-1), fx["mlp_gb"][l].astype(np.float64).reshape(-1)
        assert np.max(np.abs(a - o)) <= 4e-5 * np.max(np.abs(o)) * np.sqrt(X.shape[0]), l
    # C2 / C1
    rx = fx["ros_x"]
    _, _, grads, _ = wl.rosenbrock_batched_gradients(ctx, [ctx.array(rx[i]) for i in range(rx.shape[0])])
    got = np.stack([to_np(c)[:, 0, 0, 0] for c in grads])
    assert np.array_equal(got, fx["ros_g"])
