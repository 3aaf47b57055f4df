import os, sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import gad_b200 as gb
from helpers import to_np
ctx = gb.Context(0)
ev = gb.Eval(ctx)
def run(a, b, t1, t2):
    return to_np(ev.matmul(ev.constant(a), ev.constant(b), gb.MatProp(t1), gb.MatProp(t2)))[:, :, 0, 0].astype(np.float64)
rng = np.random.default_rng(0)
print("variant", os.environ.get("GADCU_GEMM_PAIR"), flush=True)
for (M, N, K) in [(256, 256, 32), (256, 256, 128), (512, 768, 256), (1024, 1024, 1024), (2048, 4096, 512), (4864, 4096, 64)]:
    out = []
    for t1 in (False, True):
        for t2 in (False, True):
            sa, sb = ((K, M) if t1 else (M, K)), ((N, K) if t2 else (K, N))
            a = rng.integers(-4, 5, sa).astype(np.float32); b = rng.integers(-4, 5, sb).astype(np.float32)
            got = run(a, b, t1, t2)
            opA, opB = (a.T if t1 else a).astype(np.float64), (b.T if t2 else b).astype(np.float64)
            out.append(float(np.abs(got - opA @ opB).max()))
    print((M, N, K), "maxerr NN,NT,TN,TT:", out, flush=True)
