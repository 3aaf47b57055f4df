import sys, time, numpy as np, cProfile, pstats
sys.path.insert(0, '.')
import gad_b200 as gb
from gad_b200 import workloads as wl
ctx = gb.Context(0)
B, D = 1024, 4096
X = ctx.random((B, D), 5, normal=True, a=0.0, b=1.0); T = ctx.random((B, D), 9, normal=True, a=0.0, b=1.0)
Ws = [ctx.random((D, D), 6 + l, normal=True, a=0.0, b=1.0 / np.sqrt(D)) for l in range(3)]
bs = [ctx.array(np.full((1, D), 0.01, np.float32)) for _ in range(3)]
keep = {}
def step(): keep["o"] = wl.mlp_step(ctx, X, Ws, bs, T)
for _ in range(5): step()
ctx.synchronize()
t0 = time.perf_counter()
for _ in range(50): step()
t1 = time.perf_counter(); ctx.synchronize(); t2 = time.perf_counter()
print("host issue ms/step", (t1 - t0) / 50 * 1e3, "total ms/step", (t2 - t0) / 50 * 1e3)
def loop():
    for _ in range(50): step()
    ctx.synchronize()
cProfile.run("loop()", "/tmp/prof")
pstats.Stats("/tmp/prof").sort_stats("tottime").print_stats(14)
