import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import __graft_entry__ as ge
ge.smoke()
import gad_b200 as gb
from gad_b200 import workloads as wl
ctx = gb.Context(0); ev = gb.Eval(ctx)
rng = np.random.default_rng(0)
for (M, N, K, t1, t2) in [(129, 65, 7, False, False), (200, 136, 52, True, True), (64, 64, 4, False, True)]:
    a = rng.standard_normal((K, M) if t1 else (M, K)); b = rng.standard_normal((N, K) if t2 else (K, N))
    r = ev.matmul(ev.constant(a), ev.constant(b), gb.MatProp(t1), gb.MatProp(t2)).data().numpy()
for (M, N, K) in [(256, 256, 32), (512, 256, 96)]:
    a = rng.standard_normal((M, K)).astype(np.float32); b = rng.standard_normal((K, N)).astype(np.float32)
    for t1, t2 in [(False, False), (True, False), (False, True)]:
        aa = a.T.copy() if t1 else a; bb = b.T.copy() if t2 else b
        r = ev.matmul(ev.constant(aa), ev.constant(bb), gb.MatProp(t1), gb.MatProp(t2)).data().numpy()
ctx.set_lazy(True, 2)
x, y = rng.uniform(-1, 1, 1003).astype(np.float32), rng.uniform(0.5, 2, 1003).astype(np.float32)
z, gx, gy = wl.c3_step(ctx, ctx.array(x), ctx.array(y)); gx.numpy()
xs = -1.2 + 2.4 * rng.random((8, 33))
g, f, grads, store = wl.rosenbrock_batched_gradients(ctx, [ctx.array(xs[i]) for i in range(8)])
store.rerun([ctx.array(xs[i]) for i in range(8)])[0].numpy()
print("sanitize script done")
