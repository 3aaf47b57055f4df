"""GPU: fused GEMM epilogues (gadcu_ctx_set_gemm_epilogue) against the same tape with every product launched where it is
recorded and its elementwise tail run as separate kernels. The fused tails — tanh(matmul + bias), the tanh VJP on an
incoming dA, target - (matmul + bias) — perform the tape's op sequence on the accumulator with one rounding per op, so
everything a step returns must be EQUAL, bit for bit, for aligned and ragged shapes, for products whose last wave is
split along k (the head of a split tile applies the tail, its partner stores a raw partial sum), with the TF32 operand
copies written by the epilogue feeding the next products."""
import numpy as np
import pytest

import gad_b200 as gb
from gad_b200 import workloads as wl
from helpers import to_np
import mlp_f64

pytestmark = pytest.mark.gpu


def _step(ctx, fused, X, Ws, bs, T):
    ctx.set_gemm_epilogue(2 if fused else 0)  # (2: also at shapes too small for the default to bother)
    try:
        l0 = ctx.launches()
        ctx.profile_begin()
        loss, gW, gb_ = wl.mlp_step(ctx, ctx.array(X), [ctx.array(w) for w in Ws], [ctx.array(b) for b in bs], ctx.array(T))
        out = (loss.item(), [to_np(g) for g in gW], [to_np(g) for g in gb_])
        prof = ctx.profile_end()
        return out, ctx.launches() - l0, prof
    finally:
        ctx.set_gemm_epilogue(1)


@pytest.mark.parametrize("B,D", [(512, 256), (1024, 512), (1000, 264), (2048, 1024), (4096, 1280)])
def test_fused_epilogues_equal_the_unfused_step(ctx, B, D):
    X, Ws, bs, T = mlp_f64.mlp_inputs(B, D)
    bs = [np.linspace(-0.5, 0.5, D, dtype=np.float32).reshape(1, D) * (l + 1) for l in range(3)]  # (a bias that matters)
    a, la, pa = _step(ctx, True, X, Ws, bs, T)
    b, lb, pb = _step(ctx, False, X, Ws, bs, T)
    assert a[0] == b[0]
    for x, y in zip(a[1] + a[2], b[1] + b[2]):
        assert np.array_equal(x, y)
    # the fused step launches fewer kernels: no separate bias/tanh pass, no tanh-VJP pass, fewer operand splits
    assert la < lb, (la, lb)
    assert pa["matmul"]["launches"] < pb["matmul"]["launches"] or pa["elementwise"]["launches"] < pb["elementwise"]["launches"]
    # and both are right
    want_loss, wW, wb = mlp_f64.mlp_grad_f64(X, Ws, bs, T)
    assert abs(a[0] - want_loss) <= 1e-4 * abs(want_loss)
    for l in range(3):
        assert np.abs(a[1][l][:, :, 0, 0] - wW[l]).max() <= 2e-5 * np.abs(wW[l]).max() * np.sqrt(B)


def test_fused_epilogue_with_non_square_layers_and_a_split_last_wave(ctx):
    """4096 x 2560 outputs: 160 tiles on 74 SM pairs leave a ragged last wave that is shared out along k, so some tiles
    reach C in two parts — the epilogue's tail must run exactly once per element, on the complete sum."""
    rng = np.random.default_rng(3)
    B, D0, D1 = 4096, 1024, 2560
    x = rng.standard_normal((B, D0)).astype(np.float32)
    w1 = (rng.standard_normal((D0, D1)) / np.sqrt(D0)).astype(np.float32)
    w2 = (rng.standard_normal((D1, D0)) / np.sqrt(D1)).astype(np.float32)
    b1 = rng.standard_normal((1, D1)).astype(np.float32) * 0.1
    b2 = rng.standard_normal((1, D0)).astype(np.float32) * 0.1
    t = rng.standard_normal((B, D0)).astype(np.float32)

    def run(fused):
        ctx.set_gemm_epilogue(2 if fused else 0)
        try:
            g = gb.Graph1(ctx)
            vw1, vw2, vb1, vb2 = g.variable(w1), g.variable(w2), g.variable(b1), g.variable(b2)
            h = g.tanh(g.add(g.matmul_nn(g.constant(x), vw1), g.tile_as(vb1, (B, D1))))
            h2 = g.tanh(g.add(g.matmul_nn(h, vw2), g.tile_as(vb2, (B, D0))))
            out = g.add(g.matmul_nn(h2, vw1), g.tile_as(vb1, (B, D1)))          # w1, b1 used twice: gradients accumulate
            loss = g.norm2(g.sub(g.constant(rng.standard_normal((B, D1)).astype(np.float32) * 0 + 0.25), out))
            store = g.evaluate_gradients_once(loss.gid(), 1.0)
            return loss.data().item(), [to_np(store.get(v.gid())) for v in (vw1, vw2, vb1, vb2)], to_np(h2)
        finally:
            ctx.set_gemm_epilogue(1)

    a, b = run(True), run(False)
    assert a[0] == b[0]
    assert np.array_equal(a[2], b[2])
    for x_, y_ in zip(a[1], b[1]):
        assert np.array_equal(x_, y_)


def test_a_deferred_product_can_still_be_read_on_its_own(ctx):
    """z = matmul(x, w) is recorded, not launched; reading z itself launches the plain product, and the tail computed from
    it afterwards equals the fused one."""
    rng = np.random.default_rng(5)
    B, D = 512, 256
    x, w = rng.standard_normal((B, D)).astype(np.float32), (rng.standard_normal((D, D)) / 16).astype(np.float32)
    b = rng.standard_normal((1, D)).astype(np.float32)
    ctx.set_gemm_epilogue(2)
    try:
        _deferred_product_checks(ctx, x, w, b, B, D)
    finally:
        ctx.set_gemm_epilogue(1)


def _deferred_product_checks(ctx, x, w, b, B, D):
    ev = gb.Eval(ctx)
    z = ev.matmul_nn(ev.constant(x), ev.constant(w))
    h = ev.tanh(ev.add(z, ev.tile_as(ev.constant(b), (B, D))))
    hz = to_np(h)            # fused: tanh(acc + b) in the epilogue
    zz = to_np(z)            # the product itself, launched now
    assert np.array_equal(hz[:, :, 0, 0], np.tanh((zz[:, :, 0, 0] + b).astype(np.float32)).astype(np.float32)) or \
        np.abs(hz[:, :, 0, 0] - np.tanh(zz[:, :, 0, 0].astype(np.float64) + b)).max() <= 4 * np.finfo(np.float32).eps
    want = x.astype(np.float64) @ w.astype(np.float64)
    assert np.abs(zz[:, :, 0, 0] - want).max() <= 1e-5 * np.abs(want).max()
    # an in-place write into an operand of a product that has not been launched yet forces it first
    wd = ctx.array(w)
    z2 = ev.matmul_nn(ev.constant(x), ev.constant(wd))
    wd.add_assign(ctx.array(np.ones((D, D), np.float32)))
    assert np.abs(to_np(z2)[:, :, 0, 0] - want).max() <= 1e-5 * np.abs(want).max()
