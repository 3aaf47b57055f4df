"""Driver for the ncu captures of the secondary metrics' kernels (C3 reverse sweep, C2 batched lane program).
    python profiles/run_secondary.py c3|c2
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gad_b200 as gb
from gad_b200 import workloads as wl

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
ctx = gb.Context(0)
if which == "c3":
    n = 64 * 1024 * 1024
    x = ctx.random((n,), 3, a=-1.0, b=1.0)
    y = ctx.random((n,), 4, a=0.5, b=2.0)
    for _ in range(4):
        out = wl.c3_step(ctx, x, y)   # forward kernel (lane_program #1), reduction, sweep kernel (lane_program #2)
    ctx.synchronize()
    print("z", out[0].item())
else:
    lanes, N = 1 << 20, 64
    cols = [ctx.random((lanes,), 200 + i, dtype=gb.F64, a=-1.2, b=1.2) for i in range(N)]
    for _ in range(3):
        r = wl.rosenbrock_batched_gradients(ctx, cols)
    ctx.synchronize()
    print("stats", r[3].program_stats())
