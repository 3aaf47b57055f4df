import sys, time, ctypes as C
sys.path.insert(0, '.')
import gad_b200 as gb
from gad_b200 import workloads as wl
ctx = gb.Context(0)
lanes, N = 1 << 20, 64
cols = [ctx.random((lanes,), 200 + i, dtype=gb.F64, a=-1.2, b=1.2) for i in range(N)]
g, f, grads, store = wl.rosenbrock_batched_gradients(ctx, cols)
cols2 = [ctx.random((lanes,), 300 + i, dtype=gb.F64, a=-1.2, b=1.2) for i in range(N)]
store.rerun(cols2); ctx.synchronize()
import cProfile, pstats
def loop():
    for _ in range(20):
        r = store.rerun(cols2)
    ctx.synchronize()
t0 = time.perf_counter(); loop(); print("wall per rerun ms", (time.perf_counter() - t0) / 20 * 1e3)
# raw C call
n = len(cols2)
arr = (C.c_void_p * n)(*[c.handle for c in cols2]); out = (C.c_void_p * n)()
ctx.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    ctx.lib.gadcu_batched_rerun(store.handle, arr, n, out)
    for h in out: ctx.lib.gadcu_array_release(C.c_void_p(h))
t1 = time.perf_counter(); ctx.synchronize(); t2 = time.perf_counter()
print("raw C call issue ms", (t1 - t0) / 20 * 1e3, "incl sync", (t2 - t0) / 20 * 1e3)
cProfile.run("loop()", "/tmp/prof"); pstats.Stats("/tmp/prof").sort_stats("cumtime").print_stats(8)
