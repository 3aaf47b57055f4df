import sys, numpy as np
sys.path.insert(0, '.')
import gad_b200 as gb
from gad_b200 import workloads as wl
ctx = gb.Context(0)
B, D = 2048, 1024
X = ctx.random((B, D), 5, normal=True, a=0.0, b=1.0); T = ctx.random((B, D), 9, normal=True, a=0.0, b=1.0)
Ws = [ctx.random((D, D), 6 + l, normal=True, a=0.0, b=1.0 / np.sqrt(D)) for l in range(3)]
bs = [ctx.array(np.full((1, D), 0.01, np.float32)) for _ in range(3)]
n = 1 << 22
x = ctx.random((n,), 3, a=-1.0, b=1.0); y = ctx.random((n,), 4, a=0.5, b=2.0)
cols = [ctx.random((1 << 16,), 200 + i, dtype=gb.F64, a=-1.2, b=1.2) for i in range(16)]
def one():
    out = wl.mlp_step(ctx, X, Ws, bs, T); out[0].item()
    z, gx, gy = wl.c3_step(ctx, x, y); z.item()
    g, f, grads, store = wl.rosenbrock_batched_gradients(ctx, cols); store.rerun(cols)
for i in range(5): one()
ctx.synchronize(); s0 = ctx.stats()
for i in range(100): one()
ctx.synchronize(); s1 = ctx.stats()
print("after warm-up:", s0); print("after 100 more: ", s1)
assert s1["bytes_reserved"] == s0["bytes_reserved"], "the pool grew"
assert s1["bytes_in_use"] == s0["bytes_in_use"], "live bytes grew"
print("no growth")
