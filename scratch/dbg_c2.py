import os, sys, time, numpy as np
os.environ.setdefault("GADCU_JIT_DUMP", "gpurun_out/jit_dump_c2.cu")
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import gad_b200 as gb
from gad_b200 import workloads as wl
import oracle_py
ctx = gb.Context(0)
rng = np.random.default_rng(0)
xs = -1.2 + 2.4 * rng.random((64, 5000))
t0 = time.time()
_, _, grads, store = wl.rosenbrock_batched_gradients(ctx, [ctx.array(xs[i]) for i in range(64)])
got = np.stack([g.numpy()[:, 0, 0, 0] for g in grads]); t1 = time.time()
_, og, _ = oracle_py.rosenbrock_batch(xs, nthreads=4)
print("batched equal to scalar tape:", np.array_equal(got, og), "first call s", t1 - t0, store.program_stats(), flush=True)
lanes = 1 << 20
cols = [ctx.random((lanes,), 200 + i, dtype=gb.F64, a=-1.2, b=1.2) for i in range(64)]
for _ in range(2): r = wl.rosenbrock_batched_gradients(ctx, cols)
ctx.synchronize(); ctx.profile_begin(); t0 = time.time()
for _ in range(5): r = wl.rosenbrock_batched_gradients(ctx, cols)
ctx.synchronize(); wall = (time.time() - t0) / 5; prof = ctx.profile_end()
k = prof["tile_transpose"]["ms"] / prof["tile_transpose"]["launches"]
print("batched 1M lanes: kernel ms", k, "GB/s", lanes * 1024 / k / 1e6, "wall ms", wall * 1e3, flush=True)
