"""Multi-GPU check of the data-parallel sweeps, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/dp_check_multigpu.py [B D]
(launched by tests/test_gpu_multigpu.py when the box has >= 2 GPUs). Every rank asserts; exit code 0 = all checks passed.

  C4  the fused peer-memory exchange == one grouped NCCL all-reduce of the same local gradients (bit for bit at 2 ranks
      unless a GEMM split its k range, to f32 rounding of an N-term sum beyond) == the single-GPU full batch; a second
      fused step reproduces the first exactly (the owner adds the partial tiles in rank order);
  aliasing  y = add(w1, w2): both variables receive the SAME gradient array, which must be summed over the ranks once,
      and the caller's seed must not be overwritten; matmul(W, W): both products feed one variable, none is "its whole
      gradient" (ADVICE r01);
  C5  Hessian-vector products with the overlapped per-output all-reduce == the grouped one == the single-GPU full batch.
"""
import os
import sys

import numpy as np

os.environ["NCCL_DEBUG"] = "WARN"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import gad_b200 as gb
from gad_b200 import workloads as wl

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
B, D = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) >= 3 else (2048, 1024)
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = gb.Context(lr)
obj = [gb.Comm.unique_id() if rank == 0 else None]
dist.broadcast_object_list(obj, src=0)
comm = gb.Comm(ctx, world, rank, obj[0])
rows = B // world
rng = np.random.default_rng(5)
X = rng.standard_normal((B, D)).astype(np.float32)
T = rng.standard_normal((B, D)).astype(np.float32)
Ws = [(rng.standard_normal((D, D)) / np.sqrt(D)).astype(np.float32) for _ in range(3)]
bs = [np.full((1, D), 0.01, np.float32) for _ in range(3)]
sl = slice(rank * rows, (rank + 1) * rows)
dW, db = [ctx.array(w) for w in Ws], [ctx.array(b) for b in bs]
dXs, dTs = ctx.array(X[sl]), ctx.array(T[sl])


def rel(x, y):
    return float(np.abs(x.astype(np.float64) - y.astype(np.float64)).max() / np.abs(y).max())


# ---- C4 ------------------------------------------------------------------------------------------------------------
def run(overlap):
    loss, gW, gb_ = wl.mlp_step(ctx, dXs, dW, db, dTs, comm, overlap)
    return loss.item(), [g.numpy() for g in gW], [g.numpy() for g in gb_]


a, b = run(True), run(False)
c = run(True)
assert c[0] == a[0] and all(np.array_equal(x, y) for x, y in zip(a[1] + a[2], c[1] + c[2])), "second fused step differs"
full = wl.mlp_step(ctx, ctx.array(X), dW, db, ctx.array(T))
fl, fW, fb = full[0].item(), [g.numpy() for g in full[1]], [g.numpy() for g in full[2]]
exact = a[0] == b[0] and all(np.array_equal(x, y) for x, y in zip(a[1] + a[2], b[1] + b[2]))
close = max(rel(x, y) for x, y in zip(a[1] + a[2], b[1] + b[2]))
# (not bit for bit when the two paths cut a product differently: the exchange splits the sweep's last dW GEMM into column
# chunks whose ragged last wave is shared out along k, the plain sweep does not — same products, other association)
assert close <= 5e-5, (exact, close)
vs_full = max(rel(x, y) for x, y in zip(a[1] + a[2], fW + fb))
assert vs_full <= 2e-5 * np.sqrt(B) / 8, vs_full
assert abs(a[0] - fl) <= 1e-5 * abs(fl)
msg = [f"C4 B={B} D={D}: fused==nccl exact={exact} max_rel={close:.1e}; vs single-GPU full batch {vs_full:.1e}"]

# ---- aliasing: one gradient array under two ids; the seed; matmul(W, W) ----------------------------------------------
n = 1 << 16
w1, w2 = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
seed_np = (rng.standard_normal(n) * (rank + 1)).astype(np.float32)  # per-rank seeds: the sum over ranks is not n x one of them
g = gb.Graph1(ctx)
v1, v2 = g.variable(ctx.array(w1)), g.variable(ctx.array(w2))
y = g.add(v1, v2)
seed = ctx.array(seed_np)
store = g.evaluate_gradients_once(y.gid(), seed, comm)
all_seeds = [torch.zeros(n, device="cuda") for _ in range(world)]
dist.all_gather(all_seeds, torch.from_numpy(seed_np).cuda())
want = sum(s.double() for s in all_seeds).cpu().numpy()
for v in (v1, v2):
    got = store.get(v.gid()).numpy().reshape(-1)
    assert rel(got, want) <= 1e-6, ("add(w1, w2) gradient summed wrongly", rel(got, want))
assert np.array_equal(seed.numpy().reshape(-1), seed_np), "the caller's seed was overwritten by the exchange"
msg.append("add(w1,w2): one reduction shared by both ids, seed intact")

Dm = 512
Wm = (rng.standard_normal((Dm, Dm)) / np.sqrt(Dm)).astype(np.float32)
Gm = rng.standard_normal((Dm, Dm)).astype(np.float32) * np.float32(rank + 1)
g = gb.Graph1(ctx)
vw = g.variable(ctx.array(Wm))
y = g.matmul_nn(vw, vw)
store = g.evaluate_gradients_once(y.gid(), ctx.array(Gm), comm)
tot = sum(range(1, world + 1))
G64 = Gm.astype(np.float64) / (rank + 1) * tot
want = G64 @ Wm.astype(np.float64).T + Wm.astype(np.float64).T @ G64
got = store.get(vw.gid()).numpy()[:, :, 0, 0]
assert rel(got, want) <= 1e-5, ("matmul(W, W) gradient", rel(got, want))
msg.append("matmul(W,W): both contributions summed over ranks")

# ---- C5 ------------------------------------------------------------------------------------------------------------
Bh, Dh = 512, 256
Xh, Th = rng.standard_normal((Bh, Dh)).astype(np.float32), rng.standard_normal((Bh, Dh)).astype(np.float32)
Wh = [(rng.standard_normal((Dh, Dh)) / np.sqrt(Dh)).astype(np.float32) for _ in range(3)]
bh = [np.full((1, Dh), 0.01, np.float32) for _ in range(3)]
Vh = [rng.standard_normal((Dh, Dh)).astype(np.float32) for _ in range(3)]
rh = Bh // world
slh = slice(rank * rh, (rank + 1) * rh)
args = ([ctx.array(w) for w in Wh], [ctx.array(x) for x in bh])
dV = [ctx.array(v) for v in Vh]
ho = wl.mlp_hvp(ctx, ctx.array(Xh[slh]), args[0], args[1], ctx.array(Th[slh]), dV, comm, True)
hg = wl.mlp_hvp(ctx, ctx.array(Xh[slh]), args[0], args[1], ctx.array(Th[slh]), dV, comm, False)
hf = wl.mlp_hvp(ctx, ctx.array(Xh), args[0], args[1], ctx.array(Th), dV)
for l in range(3):
    o, gg, f = ho[0][l].numpy(), hg[0][l].numpy(), hf[0][l].numpy()
    assert rel(o, gg) <= 1e-5 and rel(o, f) <= 1e-4, (l, rel(o, gg), rel(o, f))
assert abs(ho[1].item() - hf[1].item()) <= 1e-5 * abs(hf[1].item())
msg.append("C5: overlapped == grouped == single-GPU full batch")
ctx.synchronize()
dist.barrier()
print(f"rank {rank}/{world}: " + "; ".join(msg), flush=True)
dist.destroy_process_group()
