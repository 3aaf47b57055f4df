import os, sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import gad_b200 as gb
from helpers import to_np
ctx = gb.Context(0)
ev = gb.Eval(ctx)
def run(a, b, t1, t2):
    return to_np(ev.matmul(ev.constant(a), ev.constant(b), gb.MatProp(t1), gb.MatProp(t2)))[:, :, 0, 0].astype(np.float64)
rng = np.random.default_rng(0)
def sweep(M, N, K):
    out = []
    for t1 in (False, True):
        for t2 in (False, True):
            sa, sb = ((K, M) if t1 else (M, K)), ((N, K) if t2 else (K, N))
            a = rng.integers(-4, 5, sa).astype(np.float32); b = rng.integers(-4, 5, sb).astype(np.float32)
            got = run(a, b, t1, t2)
            opA, opB = (a.T if t1 else a).astype(np.float64), (b.T if t2 else b).astype(np.float64)
            out.append(float(np.abs(got - opA @ opB).max()))
    return out
variants = [None, "4096,512,1024,1,4", "512,4096,1024,1,4", "4096,1024,1024,2,3", "1024,4096,1024,2,3", "4096,512,1024,1,3", "4096,1024,1024,2,4"]
for v in variants:
    if v is None: os.environ.pop("GADCU_TC_MN_DEBUG", None)
    else: os.environ["GADCU_TC_MN_DEBUG"] = v
    try:
        print("variant", v, "maxerr NN,NT,TN,TT (128,128,32):", sweep(128, 128, 32), " (256,256,128):", sweep(256, 256, 128), flush=True)
    except Exception as e:
        print("variant", v, "failed", e, flush=True)
