"""GPU: value semantics of arrays under in-place entry points.

af::Array is a copy-on-write value: `*self += other` (WeightOps::add_assign, gad/src/net.rs:171-175) rebinds one handle;
clones, weight snapshots and values a tape captured keep their contents. The deferred elementwise regions must see the
same thing: a value recorded before a write evaluates to the pre-write contents whenever it is read, and an op recorded
after the write sees the new contents.
"""
import numpy as np
import pytest

import gad_b200 as gb
from helpers import rand, to_np

pytestmark = pytest.mark.gpu

N = 1 << 16  # above the lazy threshold: elementwise ops are deferred


def test_add_assign_is_copy_on_write(ctx):
    w0 = rand((300, 200), 1, -1, 1)
    d = rand((300, 200), 2, -1, 1)
    w = ctx.array(w0)
    snap = w.clone()                 # get_weights(): net.rs:100-130 clones
    g = gb.Graph1(ctx)
    v = g.variable(w)                # the tape takes its own clone (core.rs:162-170 takes the array by value)
    w.add_assign(ctx.array(d))
    assert np.array_equal(to_np(w)[:, :, 0, 0], w0 + d)
    assert np.array_equal(to_np(snap)[:, :, 0, 0], w0), "a clone changed under add_assign"
    assert np.array_equal(to_np(v.data())[:, :, 0, 0], w0), "the tape's forward value changed under add_assign"
    # sole holder: updated in place (same device buffer)
    u = ctx.array(w0)
    p = u.device_ptr()
    u.add_assign(ctx.array(d))
    assert u.device_ptr() == p and np.array_equal(to_np(u)[:, :, 0, 0], w0 + d)
    # shared: the handle moves to a fresh buffer, the clone keeps the old one
    u2 = u.clone()
    u.add_assign(ctx.array(d))
    assert u.device_ptr() != p and u2.device_ptr() == p
    assert np.array_equal(to_np(u2)[:, :, 0, 0], w0 + d) and np.array_equal(to_np(u)[:, :, 0, 0], (w0 + d) + d)


def test_weight_snapshot_survives_a_gradient_step(ctx):
    from gad_b200 import net as gnet
    rng = np.random.default_rng(3)
    w0 = rng.standard_normal((3, 3)).astype(np.float32)
    layer = gnet.WeightData(ctx.array(w0))
    snap = layer.get_weights()
    layer.update_weights(ctx.array(np.ones((3, 3), np.float32)))
    assert np.array_equal(to_np(snap)[:, :, 0, 0], w0), "get_weights() snapshot followed the update"
    assert np.array_equal(to_np(layer.get_weights())[:, :, 0, 0], w0 + 1)


def test_deferred_value_recorded_before_a_write_keeps_the_old_contents(ctx):
    x0 = rand((N,), 5, -1, 1)
    d = rand((N,), 6, -1, 1)
    ev = gb.Eval(ctx)
    w = ctx.array(x0)
    e_old = ev.exp(ev.constant(w))               # deferred
    w.add_assign(ctx.array(d))                   # sole holder? no: the region and the constant hold it -> rebind
    e_new = ev.exp(ev.constant(w))               # recorded after the write: must not be value-numbered to e_old
    want_old, want_new = np.exp(x0), np.exp(x0 + d)
    assert np.allclose(to_np(e_new).reshape(-1), want_new, rtol=1e-6)
    assert np.allclose(to_np(e_old).reshape(-1), want_old, rtol=1e-6)
    assert not np.array_equal(to_np(e_new), to_np(e_old))


def test_in_place_upload_between_deferred_ops(ctx):
    """An eval-only loop never opens a new tape: upload a new batch into the same buffer, run elementwise ops on it,
    keep the outputs."""
    import torch
    ev = gb.Eval(ctx)
    buf = ctx.empty((N,))
    hold = []
    wants = []
    for i in range(3):
        h = torch.from_numpy(rand((N,), 20 + i, -1, 1)).pin_memory()
        buf.upload(h.numpy())
        ctx.synchronize()
        y = ev.addc(ev.exp(ev.constant(buf)), 1.0)   # same structure every time: value numbering must not match stale registers
        hold.append(y)
        wants.append(np.exp(h.numpy()) + 1.0)
    for y, want in zip(hold, wants):
        assert np.allclose(to_np(y).reshape(-1), want, rtol=1e-6)


def test_deferred_handle_survives_a_new_tape_and_a_later_write(ctx):
    """a deferred value held across a new Graph1 (its program has left the table of open programs) is
    still computed from the pre-write contents when its leaf buffer is overwritten afterwards."""
    x0 = rand((N,), 8, -1, 1)
    buf = ctx.array(x0)
    g1 = gb.Graph1(ctx)
    y = g1.mulc(g1.exp(g1.constant(buf)), 3.0)     # deferred, in tape 1's region
    g2 = gb.Graph1(ctx)                             # a new tape: new epoch, new open programs
    z = g2.sin(g2.constant(ctx.array(rand((N,), 9, -1, 1))))
    import torch
    h = torch.from_numpy(rand((N,), 10, -1, 1)).pin_memory()
    buf.upload(h.numpy())                           # in-place write into y's leaf
    ctx.synchronize()
    assert np.allclose(to_np(y.data()).reshape(-1), np.exp(x0) * 3.0, rtol=1e-6)
    assert np.all(np.isfinite(to_np(z.data())))


def test_value_above_a_forced_value_is_not_recomputed_from_an_overwritten_leaf(ctx):
    """a = exp(w) is forced (it has a column), b = a + 1 is not; w is overwritten; a's handle goes away (its column
    returns to the pool); b must still be exp(w_old) + 1."""
    x0 = rand((N,), 13, -1, 1)
    ev = gb.Eval(ctx)
    w = ctx.array(x0)
    a = ev.exp(ev.constant(w))
    _ = to_np(a)                                   # forces a
    b = ev.addc(a, 1.0)
    import torch
    h = torch.from_numpy(rand((N,), 14, -1, 1)).pin_memory()
    w.upload(h.numpy())
    ctx.synchronize()
    del a
    assert np.allclose(to_np(b).reshape(-1), np.exp(x0) + 1.0, rtol=1e-6)
