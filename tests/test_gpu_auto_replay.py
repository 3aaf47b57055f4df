"""GPU: automatic CUDA-graph replay keyed by tape structure (gadcu_ctx_set_auto_replay). The tape is rebuilt every step, as
gad does (net_ext.rs:52); what a step issues is recorded and run as one graph, laid over the executable graph the last
step with the same kernel sequence left behind. Results must equal eager issue bit for bit, steps after the first must
reuse a graph (no re-instantiation), and every point where the host needs results must see them."""
import numpy as np
import pytest

import gad_b200 as gb
from gad_b200 import workloads as wl
from helpers import rand, to_np
import mlp_f64

pytestmark = pytest.mark.gpu


@pytest.fixture
def auto(ctx):
    ctx.synchronize()
    ctx.set_auto_replay(True)
    yield ctx
    ctx.set_auto_replay(False)


def test_c3_steps_replay_and_equal_eager(ctx):
    n = 100_003
    x, y = ctx.array(rand((n,), 3, -1, 1)), ctx.array(rand((n,), 4, 0.5, 2.0))
    z0, gx0, gy0 = wl.c3_step(ctx, x, y)
    want = (z0.item(), to_np(gx0), to_np(gy0))
    ctx.set_auto_replay(True)
    try:
        before = ctx.replay_stats()
        for _ in range(6):
            z, gx, gy = wl.c3_step(ctx, x, y)   # a fresh Graph1 every time
            assert z.item() == want[0] and np.array_equal(to_np(gx), want[1]) and np.array_equal(to_np(gy), want[2])
        after = ctx.replay_stats()
    finally:
        ctx.set_auto_replay(False)
    segments = after["segments"] - before["segments"]
    assert segments >= 6
    # only the first step's segments had to instantiate a graph; the rest updated one in place
    assert after["instantiated"] - before["instantiated"] <= segments // 6 + 2
    assert after["reused"] - before["reused"] >= segments - (segments // 6 + 2)


@pytest.mark.parametrize("B,D", [(512, 256), (100, 24)])
def test_mlp_steps_replay_and_equal_eager(ctx, B, D):
    X, Ws, bs, T = mlp_f64.mlp_inputs(B, D)
    dX, dT = ctx.array(X), ctx.array(T)
    dW, db = [ctx.array(w) for w in Ws], [ctx.array(b) for b in bs]
    l0, gW0, gb0 = wl.mlp_step(ctx, dX, dW, db, dT)
    want = (l0.item(), [to_np(g) for g in gW0], [to_np(g) for g in gb0])
    ctx.set_auto_replay(True)
    try:
        before = ctx.replay_stats()
        for _ in range(5):
            l, gW, gb_ = wl.mlp_step(ctx, dX, dW, db, dT)
            assert l.item() == want[0]
            for a, b in zip([to_np(g) for g in gW + gb_], want[1] + want[2]):
                assert np.array_equal(a, b)
        after = ctx.replay_stats()
        assert after["reused"] - before["reused"] >= 4
        # new contents in the same buffers: the replayed graphs read what is there now
        import torch
        h = torch.from_numpy(np.ascontiguousarray((2 * X).reshape(-1, order="F"))).pin_memory()
        dX.upload(h.numpy())
        l2, gW2, _ = wl.mlp_step(ctx, dX, dW, db, dT)
        got = (l2.item(), to_np(gW2[0]))
    finally:
        ctx.set_auto_replay(False)
    e2, eW2, _ = wl.mlp_step(ctx, dX, dW, db, dT)
    assert got[0] == e2.item() and np.array_equal(got[1], to_np(eW2[0]))


def test_eval_only_chains_reads_and_weight_updates(auto):
    ctx = auto
    ev = gb.Eval(ctx)
    w = ctx.array(rand((300, 200), 1, -1, 1))
    d = ctx.array(rand((300, 200), 2, -1, 1))
    want_w = to_np(w).copy()
    want_d = to_np(d)
    for i in range(4):
        y = ev.mul(ev.exp(ev.constant(w)), ev.constant(d))      # small arrays: eager kernels, recorded
        s = ev.as_scalar(ev.sum_as(y, (1, 1)))
        got = s.data().item()                                     # a host read ends the segment
        want = float(np.sum((np.exp(want_w.astype(np.float32)) * want_d).astype(np.float32), dtype=np.float64))
        assert abs(got - want) <= 1e-4 * abs(want)
        w.add_assign(d)                                           # WeightOps in place, recorded too
        want_w = (want_w + want_d).astype(np.float32)
    assert np.array_equal(to_np(w), want_w)
    stats = ctx.replay_stats()
    assert stats["segments"] > 0 and stats["reused"] > 0


def test_batched_tapes_under_auto_replay(auto):
    ctx = auto
    rng = np.random.default_rng(2)
    xs = -1.2 + 2.4 * rng.random((8, 4096))
    cols = [ctx.array(xs[i]) for i in range(8)]
    outs = []
    for _ in range(3):
        _, _, grads, _ = wl.rosenbrock_batched_gradients(ctx, cols)
        outs.append(np.stack([to_np(g)[:, 0, 0, 0] for g in grads]))
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    # closed form of d/dx_0: -400 x_0 (x_1 - x_0^2) - 2 (1 - x_0)
    want0 = -400 * xs[0] * (xs[1] - xs[0] ** 2) - 2 * (1 - xs[0])
    assert np.allclose(outs[0][0], want0, rtol=1e-12, atol=1e-9)
