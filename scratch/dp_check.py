import os, sys, numpy as np
os.environ["NCCL_DEBUG"] = "WARN"
sys.path.insert(0, '.')
import torch, torch.distributed as dist
import gad_b200 as gb
from gad_b200 import workloads as wl
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = gb.Context(lr)
obj = [gb.Comm.unique_id() if rank == 0 else None]
dist.broadcast_object_list(obj, src=0)
comm = gb.Comm(ctx, world, rank, obj[0])
B, D = 2048, 1024
rows = B // world
rng = np.random.default_rng(5)
X = rng.standard_normal((B, D)).astype(np.float32); T = rng.standard_normal((B, D)).astype(np.float32)
Ws = [(rng.standard_normal((D, D)) / np.sqrt(D)).astype(np.float32) for _ in range(3)]
bs = [np.full((1, D), 0.01, np.float32) for _ in range(3)]
sl = slice(rank * rows, (rank + 1) * rows)
def run(overlap):
    loss, gW, gb_ = wl.mlp_step(ctx, ctx.array(X[sl]), [ctx.array(w) for w in Ws], [ctx.array(b) for b in bs], ctx.array(T[sl]), comm, overlap)
    return loss.item(), [g.numpy() for g in gW], [g.numpy() for g in gb_]
a, b = run(True), run(False)
c = run(True)
assert c[0] == a[0] and all(np.array_equal(x, y) for x, y in zip(a[1] + a[2], c[1] + c[2])), "second fused step differs"
full = wl.mlp_step(ctx, ctx.array(X), [ctx.array(w) for w in Ws], [ctx.array(b) for b in bs], ctx.array(T))
fl, fW = full[0].item(), [g.numpy() for g in full[1]]
ok = a[0] == b[0] and all(np.array_equal(x, y) for x, y in zip(a[1] + a[2], b[1] + b[2]))
rel = max(float(np.abs(x - y).max() / np.abs(y).max()) for x, y in zip(a[1], fW))
print(f"rank {rank}: overlap == grouped: {ok}; loss {a[0]} vs full-batch {fl}; max rel diff of dW vs single-GPU full batch {rel:.2e}", flush=True)
dist.destroy_process_group()
