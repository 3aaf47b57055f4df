"""GPU: the asynchronous copies either side of a step — `upload_async` / `wait_uploads` (prefetch of the next batch on the
copy stream) and `download_async` / `wait_download` (a step's result read back without holding the host up). These are the
calls `bench.py`'s end-to-end loop makes every step."""
import numpy as np
import pytest
import torch

import gad_b200 as gb
from gad_b200 import workloads as wl
from helpers import rand

pytestmark = pytest.mark.gpu


def _pinned(shape, dtype=torch.float32):
    return torch.empty(shape, dtype=dtype).pin_memory()


def test_download_async_returns_the_values_of_the_step_that_produced_them(ctx):
    """Three steps in flight: each queues the read-back of its own result; the tickets are waited for afterwards, in any
    order, and every host buffer holds its own step's value (== the synchronous read)."""
    n = 200_000
    xs = [rand((n,), 30 + i, -1.0, 1.0) for i in range(3)]
    hosts = [_pinned((1,)) for _ in range(3)]
    full = _pinned((n,))
    tickets, want = [], []
    ev = gb.Eval(ctx)
    for i in range(3):
        d = ev.mul(ev.exp(ev.constant(ctx.array(xs[i]))), ev.constant(ctx.array(xs[i])))  # a deferred chain: forced by the download
        s = ev.sum_as(d, (1,))
        tickets.append(s.data().download_async(hosts[i].numpy()))
        want.append(s.data().numpy().reshape(-1)[0])
        if i == 2:
            t_full = d.data().download_async(full.numpy())
    assert tickets == sorted(tickets) and len(set(tickets)) == 3
    ctx.wait_download(tickets[1])
    assert hosts[0].numpy()[0] == want[0] and hosts[1].numpy()[0] == want[1]  # (tickets complete in order)
    ctx.wait_download(t_full)
    assert hosts[2].numpy()[0] == want[2]
    assert np.array_equal(full.numpy(), d.data().numpy().reshape(-1))
    ctx.wait_download(tickets[0])  # retired already: returns at once
    with pytest.raises(gb.InvalidError):
        ctx.wait_download(t_full + 1000)
    with pytest.raises(ValueError):
        d.data().download_async(np.empty(3, np.float32))


def test_pipelined_step_loop_matches_the_synchronous_one(ctx):
    """bench.py's end-to-end loop in small: step i's inputs were prefetched during step i-1, its loss is read one step later.
    Losses equal those of the plain loop (upload, step, item()) bit for bit, for every step."""
    B, D, steps = 256, 128, 6
    rng = np.random.default_rng(4)
    Ws = [ctx.array((rng.standard_normal((D, D)) / np.sqrt(D)).astype(np.float32)) for _ in range(3)]
    bs = [ctx.array(np.full((1, D), 0.01, np.float32)) for _ in range(3)]
    batches = [(rng.standard_normal((B, D)).astype(np.float32), rng.standard_normal((B, D)).astype(np.float32)) for _ in range(steps)]
    plain = [wl.mlp_step(ctx, ctx.array(x), Ws, bs, ctx.array(t))[0].item() for x, t in batches]

    hx, ht = [_pinned((B * D,)) for _ in range(2)], [_pinned((B * D,)) for _ in range(2)]
    bufs = [(ctx.empty((B, D)), ctx.empty((B, D))) for _ in range(2)]
    hloss = [_pinned((1,)) for _ in range(2)]

    def prefetch(i):
        x, t = batches[i]
        hx[i % 2].numpy()[:] = x.reshape(-1, order="F")
        ht[i % 2].numpy()[:] = t.reshape(-1, order="F")
        bufs[i % 2][0].upload_async(hx[i % 2].numpy())
        bufs[i % 2][1].upload_async(ht[i % 2].numpy())

    got, pending = [], None
    prefetch(0)
    for i in range(steps):
        ctx.wait_uploads()
        if pending is not None:  # the host buffers of step i-1 are about to be reused by prefetch(i+1): its loss first
            ctx.wait_download(pending)
            got.append(float(hloss[(i - 1) % 2].numpy()[0]))
        if i + 1 < steps:
            prefetch(i + 1)
        loss = wl.mlp_step(ctx, bufs[i % 2][0], Ws, bs, bufs[i % 2][1])[0]
        pending = loss.download_async(hloss[i % 2].numpy())
    ctx.wait_download(pending)
    got.append(float(hloss[(steps - 1) % 2].numpy()[0]))
    assert got == [float(np.float32(v)) for v in plain]
