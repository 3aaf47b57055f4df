import os, sys, time, numpy as np
os.environ.setdefault("GADCU_JIT_DUMP", "gpurun_out/jit_dump.cu")
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import gad_b200 as gb
from gad_b200 import workloads as wl
import oracle_py
ctx = gb.Context(0)
rng = np.random.default_rng(0)

def c3(n, fuse):
    ctx.set_fusion(fuse)
    x = np.random.default_rng(1).uniform(-1, 1, n).astype(np.float32)
    y = np.random.default_rng(2).uniform(0.5, 2, n).astype(np.float32)
    z, gx, gy = wl.c3_step(ctx, ctx.array(x), ctx.array(y))
    return z.item(), gx.numpy().reshape(-1), gy.numpy().reshape(-1)

for n in (100003, 1 << 20):
    l0 = ctx.launches(); a = c3(n, True); l1 = ctx.launches(); b = c3(n, False); l2 = ctx.launches()
    print("C3 n", n, "lazy launches", l1 - l0, "unfused launches", l2 - l1, "z", a[0], b[0], "gx equal", np.array_equal(a[1], b[1]), "gy equal", np.array_equal(a[2], b[2]),
          "max|dgx|", np.abs(a[1] - b[1]).max(), flush=True)

def mlp(B, D, fuse):
    ctx.set_fusion(fuse)
    r = np.random.default_rng(5)
    X = r.standard_normal((B, D)).astype(np.float32)
    Ws = [(r.standard_normal((D, D)) / np.sqrt(D)).astype(np.float32) for _ in range(3)]
    bs = [np.full((1, D), 0.01, np.float32) for _ in range(3)]
    T = r.standard_normal((B, D)).astype(np.float32)
    loss, gW, gb_ = wl.mlp_step(ctx, ctx.array(X), [ctx.array(w) for w in Ws], [ctx.array(b) for b in bs], ctx.array(T))
    return loss.item(), [g.numpy() for g in gW], [g.numpy() for g in gb_]
for (B, D) in ((512, 256), (1024, 512)):
    l0 = ctx.launches(); a = mlp(B, D, True); l1 = ctx.launches(); b = mlp(B, D, False); l2 = ctx.launches()
    print("MLP", B, D, "lazy launches", l1 - l0, "unfused", l2 - l1, "loss", a[0], b[0], "gW equal", [np.array_equal(x, y) for x, y in zip(a[1], b[1])],
          "gb equal", [np.array_equal(x, y) for x, y in zip(a[2], b[2])], "max rel dW", [float(np.abs(x - y).max() / np.abs(y).max()) for x, y in zip(a[1], b[1])], flush=True)
ctx.set_fusion(True)

# batched tapes through the compiled lane program
xs = -1.2 + 2.4 * rng.random((64, 5000))
t0 = time.time()
_, _, grads, store = wl.rosenbrock_batched_gradients(ctx, [ctx.array(xs[i]) for i in range(64)])
got = np.stack([g.numpy()[:, 0, 0, 0] for g in grads]); t1 = time.time()
_, og, _ = oracle_py.rosenbrock_batch(xs, nthreads=4)
print("batched equal to scalar tape:", np.array_equal(got, og), "max abs diff", np.abs(got - og).max(), "first call s", t1 - t0, store.program_stats(), flush=True)

# timing: C3 sweep at 64 Mi
import torch
ext = torch.cuda.ExternalStream(ctx.stream, device=0)
n = 64 * 1024 * 1024
x = ctx.random((n,), 3, a=-1.0, b=1.0); y = ctx.random((n,), 4, a=0.5, b=2.0)
g = gb.Graph1(ctx); vx, vy, z = wl.c3_forward(g, x, y)
keep = {}
def sweep(): keep["s"] = g.evaluate_gradients(z.gid(), 1.0)
for _ in range(3): sweep()
ctx.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = ctx.launches(); e0.record(ext)
for _ in range(10): sweep()
e1.record(ext); e1.synchronize()
ms = e0.elapsed_time(e1) / 10
print("C3 sweep ms", ms, "GB/s algorithmic", 4 * n * 4 / ms / 1e6, "launches/sweep", (ctx.launches() - l0) / 10, flush=True)
def fb(): keep["r"] = wl.c3_step(ctx, x, y)
for _ in range(3): fb()
ctx.synchronize(); e0.record(ext)
for _ in range(10): fb()
e1.record(ext); e1.synchronize()
ms = e0.elapsed_time(e1) / 10
print("C3 fwd+sweep ms", ms, "GB/s (6 passes)", 6 * n * 4 / ms / 1e6, flush=True)
# batched at full size
lanes = 1 << 20
cols = [ctx.random((lanes,), 200 + i, dtype=gb.F64, a=-1.2, b=1.2) for i in range(64)]
for _ in range(2): r = wl.rosenbrock_batched_gradients(ctx, cols)
ctx.synchronize(); ctx.profile_begin(); t0 = time.time()
for _ in range(3): r = wl.rosenbrock_batched_gradients(ctx, cols)
ctx.synchronize(); wall = (time.time() - t0) / 3; prof = ctx.profile_end()
print("batched 1M lanes: wall ms", wall * 1e3, prof, flush=True)
