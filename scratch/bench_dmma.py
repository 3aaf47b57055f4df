import sys, numpy as np
sys.path.insert(0, '.')
import torch, gad_b200 as gb
ctx = gb.Context(0); ev = gb.Eval(ctx)
ext = torch.cuda.ExternalStream(ctx.stream, device=0)
for (M, N, K) in [(4096, 4096, 4096), (8192, 4096, 4096)]:
    a = ev.constant(ctx.random((M, K), 1, dtype=gb.F64, normal=True, a=0.0, b=1.0)); b = ev.constant(ctx.random((K, N), 2, dtype=gb.F64, normal=True, a=0.0, b=1.0))
    for t1, t2 in [(False, False), (False, True), (True, False)]:
        aa = a if not t1 else ev.constant(ctx.random((K, M), 3, dtype=gb.F64, normal=True, a=0.0, b=1.0))
        bb = b if not t2 else ev.constant(ctx.random((N, K), 4, dtype=gb.F64, normal=True, a=0.0, b=1.0))
        for _ in range(2): r = ev.matmul(aa, bb, gb.MatProp(t1), gb.MatProp(t2))
        ctx.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(5): r = ev.matmul(aa, bb, gb.MatProp(t1), gb.MatProp(t2))
        e1.record(ext); e1.synchronize(); ms = e0.elapsed_time(e1) / 5
        print(f"f64 GEMM {M}x{N}x{K} tA={t1} tB={t2}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
