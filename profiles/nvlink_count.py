"""NVLink bytes per data-parallel C4 step, from the driver's per-link data counters (nvidia-smi nvlink -gt d) read on GPU 0
before and after a fixed number of steps — next to the algorithmic figure of the exchange.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 profiles/nvlink_count.py [steps]
"""
import os, re, subprocess, sys
os.environ.setdefault("NCCL_DEBUG", "WARN")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
import gad_b200 as gb
from gad_b200 import workloads as wl

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
ctx = gb.Context(lr)
obj = [gb.Comm.unique_id() if rank == 0 else None]
dist.broadcast_object_list(obj, src=0)
comm = gb.Comm(ctx, world, rank, obj[0])
B, D, L = 8192, 4096, 3
rows = B // world
X = ctx.random((rows, D), 5 + 100 * rank, normal=True, a=0.0, b=1.0)
T = ctx.random((rows, D), 9 + 100 * rank, normal=True, a=0.0, b=1.0)
Ws = [ctx.random((D, D), 6 + l, normal=True, a=0.0, b=1.0 / np.sqrt(D)) for l in range(L)]
bs = [ctx.array(np.full((1, D), 0.01, np.float32)) for _ in range(L)]


def counters():
    out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", "0"], capture_output=True, text=True).stdout
    tx = sum(int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out))
    rx = sum(int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out))
    return tx, rx, out


keep = None
for _ in range(10):
    keep = wl.mlp_step(ctx, X, Ws, bs, T, comm, True)
ctx.synchronize(); dist.barrier(); torch.cuda.synchronize()
c0 = counters() if rank == 0 else None
dist.barrier()
for _ in range(steps):
    keep = wl.mlp_step(ctx, X, Ws, bs, T, comm, True)
ctx.synchronize(); dist.barrier(); torch.cuda.synchronize()
if rank == 0:
    c1 = counters()
    tx, rx = (c1[0] - c0[0]) / 1024.0 / steps, (c1[1] - c0[1]) / 1024.0 / steps
    S = D * D * 4 / 2 ** 20
    nvls = os.environ.get("GADCU_DP_NVLS", "auto")
    uses_nvls = (world >= 4) if nvls == "auto" else nvls != "0"
    alg = L * (1 + 1.0 / world) * S if uses_nvls else L * 2 * (world - 1) / world * S
    print(f"nvlink counters, GPU 0, {world} ranks, {steps} C4 steps, exchange = {'NVLS' if uses_nvls else 'tile-pushing'}: "
          f"Tx {tx:.1f} MiB/step, Rx {rx:.1f} MiB/step; algorithmic {alg:.1f} MiB per direction and step "
          f"(3 gradients of {S:.0f} MiB: {'(1 + 1/N) S each' if uses_nvls else '2 (N-1)/N S each'}) + 48 KiB of bias gradients and the loss to every peer", flush=True)
    if tx == 0 and rx == 0:
        print("raw counter output (no data counters found?):\n" + c1[2][:1500], flush=True)
dist.destroy_process_group()
