"""CPU, world size 2, gloo: the host-side logic of the data-parallel path (gad_b200/dp.py, bench.py's rank
handling), with the oracle standing in for the device kernels. The device side of the same seam is covered on
one GPU by test_mlp_sharded_rows_sum_to_the_full_batch and on 2 GPUs by scratch-free bench runs."""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from gad_b200 import dp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_partition_the_batch():
    for total in (0, 1, 7, 8, 8192, 1_048_577):
        for world in (1, 2, 3, 8):
            shards = dp.all_shards(total, world)
            assert shards[0][0] == 0 and shards[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(shards, shards[1:]))
            sizes = [b - a for a, b in shards]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert dp.even_rows(8192, 8) == 1024
    with pytest.raises(ValueError):
        dp.even_rows(10, 4)
    with pytest.raises(ValueError):
        dp.shard_bounds(10, 2, 2)


WORKER = textwrap.dedent("""
    import os, sys, json
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import oracle_py
    from gad_b200 import dp
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    B, D, L = 48, 16, 3
    rng = np.random.default_rng(5)
    X = rng.standard_normal((B, D)).astype(np.float32); T = rng.standard_normal((B, D)).astype(np.float32)
    Ws = [(rng.standard_normal((D, D)) / np.sqrt(D)).astype(np.float32) for _ in range(L)]
    bs = [np.full((1, D), 0.01, np.float32) for _ in range(L)]
    a, b = dp.shard_bounds(B, world, rank)
    loss, gW, gb, _ = oracle_py.mlp_step(X[a:b], Ws, bs, T[a:b])
    # the one exchange step: sum of the weight / bias gradients and of the loss over ranks
    flat = torch.from_numpy(np.concatenate([g.reshape(-1) for g in gW + gb] + [np.float32([loss])]))
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if rank == 0:
        floss, fW, fb, _ = oracle_py.mlp_step(X, Ws, bs, T)
        want = np.concatenate([g.reshape(-1) for g in fW + fb] + [np.float32([floss])])
        got = flat.numpy()
        scale = np.abs(want).max()
        print(json.dumps({{"max_abs_diff": float(np.abs(got - want).max()), "scale": float(scale), "rows": [a, b]}}))
    dist.barrier()
    dist.destroy_process_group()
""")


def _torchrun(args, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533"] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_sharded_gradients_sum_to_the_full_batch_over_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    res = _torchrun([str(script)])
    assert res.returncode == 0, res.stdout + res.stderr
    line = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(line) == 1, res.stdout
    out = json.loads(line[0])
    assert out["rows"] == [0, 24]
    # same sum, different association (two partial sums instead of one pass over the rows)
    assert out["max_abs_diff"] <= 1e-5 * out["scale"], out


def test_reference_arm_prints_one_line_from_rank_0_only():
    res = _torchrun(["bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"])
    assert res.returncode == 0, res.stdout + res.stderr
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, res.stdout
    out = json.loads(lines[0])
    assert out["impl"] == "reference" and out["n_gpus"] == 2 and out["unit"] == "steps/s"
    assert out["cpu_baseline"]["kind"] == "port" and out["e2e"]["h2d_bytes_per_step"] == 0
