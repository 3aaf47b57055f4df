#!/usr/bin/env python
"""bench.py — BASELINE.json's headline metric on B200, measured through the C ABI.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input: a fresh Graph1 tape, the
3-layer 4096-wide tanh MLP forward over batch 8192 (f32, column-major), square loss, the reverse sweep
(evaluate_gradients_once), and with N > 1 the NCCL all-reduce of the weight gradients — BASELINE
config C4, batch rows sharded over ranks (total work fixed: strong scaling). Rank 0 prints ONE JSON line.

  value      whole-job steps/s with inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e        the same metric through the public API with HOST buffers: every step copies that step's
             X / target rows from pinned host memory and reads the loss back
  roofline   the dominant kernel (matmul): tensor-pipe FLOP per launch / measured launch duration vs the
             measured TF32-dense peak (half the bf16 figure of MEASURED_PEAKS.json)
  secondary  the other two BASELINE metrics at N=1: reverse-sweep GB/s (C3) and batched-tape grads/s (C2)
  cpu_baseline  the oracle (a C++ port of the reference's CPU tape; the reference itself is Rust +
             arrayfire and cannot be built here) timed on the box's host cores on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_TOTAL, D, L = 8192, 4096, 3
GEMMS_PER_STEP = 8  # 3 forward + 5 backward (X is a constant: no dX)


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update({k: m[k] for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained") if k in m})
        p["source"] = "measured"
    except Exception:
        pass
    return p


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons sampled DURING the timed region: NVML in-process every 10 ms (a short
    timed region still gets tens of samples), nvidia-smi every 200 ms if NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def run_nvml(self):
        nv = self.nvml
        bits = [("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8),
                ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)]
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append([str(self.index), str(sm), str(self.max_sm), str(pw)] +
                                    ["Active" if mask & b else "Not Active" for _, b in bits])
            except Exception:
                pass
            time.sleep(0.01)

    def run(self):
        if self.nvml is not None:
            return self.run_nvml()
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        sm = [float(s[1]) for s in self.samples if len(s) > 2 and s[1].replace(".", "").isdigit()]
        mx = [float(s[2]) for s in self.samples if len(s) > 2 and s[2].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        pw = [float(s[3]) for s in self.samples if len(s) > 3 and s[3].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(self.samples),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def ncu_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores (the ONLY use of oracle/ here)
# ---------------------------------------------------------------------------------------------------
def cpu_mlp_steps_per_s(rows: int, threads: int, repeats: int = 1):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py
    rng = np.random.default_rng(5)
    X = rng.standard_normal((rows, D)).astype(np.float32)
    Ws = [(rng.standard_normal((D, D)) / np.sqrt(D)).astype(np.float32) for _ in range(L)]
    bs = [np.full((1, D), 0.01, np.float32) for _ in range(L)]
    T = rng.standard_normal((rows, D)).astype(np.float32)
    best = None
    for _ in range(repeats):
        _, _, _, secs = oracle_py.mlp_step(X, Ws, bs, T, nthreads=threads)
        best = secs if best is None else min(best, secs)
    # a full step is B_TOTAL/rows of these (every GEMM, elementwise pass and reduction is linear in rows)
    return 1.0 / (best * (B_TOTAL / rows)), best


def cpu_secondary(threads: int):
    """The restated CPU path beside the secondary metrics, each on a bounded sample (about a second of CPU work):
    C1 one scalar Rosenbrock-64 tape on one core, C2 the same tape looped over lanes on all cores, C3 the HostArray tape
    (one core: the reference's CPU arrays are single-threaded per op) with the reverse sweep timed on its own."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py
    rng = np.random.default_rng(2)
    out = {}
    x1 = -1.2 + 2.4 * rng.random((64, 1024))
    _, _, secs = oracle_py.rosenbrock_batch(x1, nthreads=1)
    out["c1"] = {"value": x1.shape[1] / secs, "unit": "grads/s", "cores": 1, "kind": "port",
                 "sample": "1024 Rosenbrock-64 f64 scalar Graph1 tapes, recorded and swept one after the other"}
    lanes = 2048 * threads
    xs = -1.2 + 2.4 * rng.random((64, lanes))
    _, _, secs = oracle_py.rosenbrock_batch(xs, nthreads=threads)
    out["c2"] = {"value": lanes / secs, "unit": "grads/s", "cores": threads, "kind": "port",
                 "sample": f"{lanes} of 1048576 lanes (the scalar tape looped over lanes, {threads} threads); grads/s needs no scaling"}
    n = 1 << 22
    x = rng.uniform(-1.0, 1.0, n).astype(np.float32)
    y = rng.uniform(0.5, 2.0, n).astype(np.float32)
    _, _, _, secs, sweep = oracle_py.c3(x, y)
    out["c3"] = {"value": 16.0 * n / sweep / 1e9, "unit": "GB/s", "cores": 1, "kind": "port", "fwd_plus_sweep_seconds": secs,
                 "sample": f"n = {n} of 67108864 elements (HostArray tape, per-node sweep); algorithmic GB/s = 16 B/element / sweep time, size-independent"}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_cores()
    rows = 512  # ~4 s per step on 16 cores
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_mlp_steps_per_s(rows, threads)
    vals, times = [], []
    for _ in range(max(1, min(args.steps, 5))):
        v, t = cpu_mlp_steps_per_s(rows, threads)
        vals.append(v)
        times.append(t)
    v = float(np.mean(vals))
    sample = f"{rows} of {B_TOTAL} batch rows at full width D={D}, L={L}; steps/s scaled by {B_TOTAL // rows}x (work is linear in rows)"
    line = {"impl": "reference", "metric": "MLP fwd+bwd steps/s", "value": v, "unit": "steps/s", "n_gpus": args.gpus,
            "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C4 3-layer tanh MLP {D}-wide, batch {B_TOTAL}, f32, Graph1 fwd + reverse sweep", "global_batch": B_TOTAL},
            "cpu_baseline": {"value": v, "unit": "steps/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference (Rust crate gad + arrayfire) cannot be built in this image; this is oracle/, a C++ port of its CPU tape"}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import gad_b200 as gb
    from gad_b200 import workloads as wl

    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = gb.Context(local_rank)
    ext = torch.cuda.ExternalStream(ctx.stream, device=local_rank)

    comm = None
    if world > 1:
        obj = [gb.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        comm = gb.Comm(ctx, world, rank, obj[0])

    from gad_b200 import dp
    rows = dp.even_rows(B_TOTAL, world)
    # synthetic data of the named shapes: X~N(0,1) seed 5, W_l~N(0,1/D) seeds 6-8, b=0.01, target~N(0,1) seed 9
    X = ctx.random((rows, D), 5 + 100 * rank, normal=True, a=0.0, b=1.0)
    T = ctx.random((rows, D), 9 + 100 * rank, normal=True, a=0.0, b=1.0)
    Ws = [ctx.random((D, D), 6 + l, normal=True, a=0.0, b=1.0 / np.sqrt(D)) for l in range(L)]
    bs = [ctx.array(np.full((1, D), 0.01, np.float32)) for _ in range(L)]

    def barrier():
        ctx.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        t_host = time.perf_counter()
        for _ in range(steps):
            fn()
        host_ms[0] = (time.perf_counter() - t_host) * 1e3 / steps  # time to ISSUE a step (the device may lag behind)
        e1.record(ext)
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        barrier()
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    keep = {}

    overlap = os.environ.get("GADCU_DP_OVERLAP", "1") == "1"  # fused GEMM -> reduce-scatter over peer memory

    def step_resident():
        keep["out"] = wl.mlp_step(ctx, X, Ws, bs, T, comm, overlap)

    for _ in range(max(3, args.warmup)):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launches()
    ctx.profile_begin()
    ms = timed(step_resident, args.steps)
    prof = ctx.profile_end()
    launches = ctx.launches() - launches0
    host_issue_ms = host_ms[0]
    clocks = sampler.summary() if rank == 0 else None
    loss_val = keep["out"][0].item()
    steps_per_s = args.steps / (ms / 1000.0)

    # ---- e2e: host buffers in, loss out, every step
    hX = torch.empty((rows * D,), dtype=torch.float32).pin_memory()
    hT = torch.empty((rows * D,), dtype=torch.float32).pin_memory()
    hX.normal_(generator=torch.Generator().manual_seed(5 + rank))
    hT.normal_(generator=torch.Generator().manual_seed(9 + rank))
    nX, nT = hX.numpy(), hT.numpy()
    # double-buffered input prefetch, as a training loop's loader does it: while step i computes, the copy stream
    # brings step i+1's X / target rows from pinned host memory. Every timed step issues exactly one such pair of
    # copies and reads its own loss back (which also orders step i+1 behind step i's reads of the buffers).
    bufs = [(ctx.empty((rows, D)), ctx.empty((rows, D))) for _ in range(2)]
    e2e_loss, e2e_i = {}, [0]

    def prefetch(i):
        bx, bt = bufs[i % 2]
        bx.upload_async(nX)
        bt.upload_async(nT)

    def step_e2e():
        i = e2e_i[0]
        ctx.wait_uploads()       # this step's inputs (copied during the previous step)
        prefetch(i + 1)          # next step's inputs, overlapping this step's kernels
        bx, bt = bufs[i % 2]
        out = wl.mlp_step(ctx, bx, Ws, bs, bt, comm, overlap)
        e2e_loss["v"] = out[0].item()  # device -> host read of the step's result
        e2e_i[0] = i + 1

    prefetch(0)
    for _ in range(3):
        step_e2e()
    e2e_steps = max(1, args.steps)
    ms_e2e = timed(step_e2e, e2e_steps)
    e2e = {"value": e2e_steps / (ms_e2e / 1000.0), "unit": "steps/s", "h2d_bytes_per_step": int(2 * rows * D * 4 * world),
           "d2h_bytes_per_step": int(4 * world), "ms_per_step": ms_e2e / e2e_steps}

    hvp_multi = None
    if world > 1 and not args.headline_only:
        hvp_multi = hvp_metric(ctx, ext, torch, wl, comm, world, rank, dist)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    pk = peaks()
    # TF32 dense = half of the measured bf16 figure. MEASURED_PEAKS.json has two: a burst (clocks at max) and a
    # sustained one (seconds under the power cap, ~1.3 GHz). The denominator is the one whose clock regime the timed
    # region was in, by the SM clocks sampled during it; both fractions are reported.
    burst_regime = bool(clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] >= 0.9 * clocks["sm_max_mhz"])
    peak_key = "bf16_tflops" if burst_regime else "bf16_tflops_sustained"
    tf32_peak = pk[peak_key] / 2.0
    gemm = prof["matmul"]
    gemm_ms = gemm["ms"] / max(1, gemm["launches"])
    logical_flop = 2.0 * rows * D * D
    tc_path = os.environ.get("GADCU_GEMM_PATH", "auto")
    mma_flop = 3.0 * logical_flop  # 3xTF32: three tensor-core products per logical product
    achieved = mma_flop / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    roofline = {"bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved / tf32_peak,
                "traffic": ncu_traffic("gemm_tf32x3_pair_kernel"), "kernel": "gemm_tf32x3_pair_kernel + its operand splits (fwd / dA / dW), 3xTF32",
                "frac_of_burst_peak": achieved / (pk["bf16_tflops"] / 2.0), "frac_of_sustained_peak": achieved / (pk["bf16_tflops_sustained"] / 2.0),
                "logical_fp32_tflops": logical_flop / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0,
                "peak_source": f"{pk['source']}: {peak_key}/2 (sampled SM clock {'at' if burst_regime else 'below'} max)", "avg_launch_ms": gemm_ms, "launches": gemm["launches"],
                "share_of_step": gemm["ms"] / ms if ms > 0 else None, "gemm_path": tc_path}

    line = {"metric": "MLP fwd+bwd steps/s", "value": steps_per_s, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C4 3-layer tanh MLP {D}-wide, batch {B_TOTAL}, f32, Graph1 fwd + reverse sweep", "global_batch": B_TOTAL,
                       "rows_per_gpu": rows, "parallelism": f"dp{world} (batch rows sharded, NCCL allreduce of weight grads" + (": dW GEMM epilogues reduce-scatter over NVLink peer memory, gather overlapped with the sweep)" if overlap and world > 1 else ")"),
                       "l2": "inputs larger than L2 (activations 128 MiB, weights 64 MiB each vs 126 MB L2)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "kernel_time_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()}, "host_issue_ms_per_step": host_issue_ms,
            "loss": loss_val}

    if hvp_multi is not None:
        line["secondary"] = {"hvp": hvp_multi}
    if world == 1 and not args.headline_only:
        try:
            line["secondary"] = secondary_metrics(ctx, ext, torch, gb, wl, pk, args)
        except Exception as e:  # the headline line must survive
            line["secondary"] = {"error": repr(e)}
        try:
            threads = host_cores()
            rows_s = 1024  # ~10 s of CPU work on 16 cores: one eighth of the batch at full width, all L layers
            v, t = cpu_mlp_steps_per_s(rows_s, threads)
            line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": threads, "kind": "port", "seconds": t,
                                    "sample": f"{rows_s} of {B_TOTAL} batch rows at full width D={D}, L={L}; scaled by {B_TOTAL // rows_s}x"}
        except Exception as e:
            line["cpu_baseline"] = {"error": repr(e)}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def secondary_metrics(ctx, ext, torch, gb, wl, pk, args):
    out = {}

    def ev_time(fn, reps):
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(reps):
            fn()
        e1.record(ext)
        e1.synchronize()
        return e0.elapsed_time(e1) / reps

    # ---- C3: reverse sweep of the 64 Mi-element elementwise tape
    n = 64 * 1024 * 1024
    x = ctx.random((n,), 3, a=-1.0, b=1.0)
    y = ctx.random((n,), 4, a=0.5, b=2.0)
    g = gb.Graph1(ctx)
    vx, vy, z = wl.c3_forward(g, x, y)
    keep = {}

    def sweep():
        keep["s"] = g.evaluate_gradients(z.gid(), 1.0)

    for _ in range(3):
        sweep()
    l0 = ctx.launches()
    ms_sweep = ev_time(sweep, 10)
    sweep_launches = (ctx.launches() - l0) // 10
    alg_bytes = 4 * n * 4  # read x, y; write dz/dx, dz/dy
    gbs = alg_bytes / (ms_sweep * 1e-3) / 1e9

    def fwd_bwd():
        keep["r"] = wl.c3_step(ctx, x, y)

    for _ in range(3):
        fwd_bwd()
    ms_fb = ev_time(fwd_bwd, 10)
    out["grad_sweep"] = {
        "metric": "grad-sweep HBM GB/s", "workload": "C3 64Mi-element exp/log/mul/add + sum, f32, reverse sweep", "value": gbs, "unit": "GB/s",
        "ms_per_sweep": ms_sweep, "launches_per_sweep": int(sweep_launches),
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                     "frac_of_8TBs": gbs / 8000.0, "traffic": ncu_traffic("lane_program_c3_sweep"), "algorithmic_bytes": alg_bytes,
                     "note": "algorithmic = 4 passes (read x,y; write gx,gy); the per-node sweep materialises more (see DESIGN.md)"},
        "fwd_plus_sweep_ms": ms_fb, "fwd_plus_sweep_gbs": 6 * n * 4 / (ms_fb * 1e-3) / 1e9}
    del keep, g, vx, vy, z, x, y

    # ---- C2: batched scalar tapes, 2^20 lanes x Rosenbrock-64, f64
    lanes, N = 1 << 20, 64
    cols = [ctx.random((lanes,), 200 + i, dtype=gb.F64, a=-1.2, b=1.2) for i in range(N)]
    res = {}

    def batched():
        res["r"] = wl.rosenbrock_batched_gradients(ctx, cols)

    batched()
    ctx.profile_begin()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        batched()
    ctx.synchronize()
    wall = (time.perf_counter() - t0) / reps
    prof = ctx.profile_end()
    k_ms = prof["tile_transpose"]["ms"] / max(1, prof["tile_transpose"]["launches"])
    stats = res["r"][3].program_stats()
    gps = lanes / (k_ms * 1e-3) if k_ms > 0 else 0.0
    # steady state of an optimiser loop: the same tapes at new points through the compiled program (no re-recording)
    store = res["r"][3]
    cols2 = [ctx.random((lanes,), 300 + i, dtype=gb.F64, a=-1.2, b=1.2) for i in range(N)]
    for _ in range(3):  # (the first calls grow the pool by a second set of output columns)
        keep_r = store.rerun(cols2)
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        keep_r = store.rerun(cols2)
    ctx.synchronize()
    rerun_wall = (time.perf_counter() - t0) / 10
    alg = 2 * N * 8  # bytes per gradient: read 64 f64, write 64 f64
    out["batched_tape"] = {
        "metric": "batched-tape grads/s", "workload": f"C2 {lanes} lanes x Rosenbrock-{N}, f64, one tape per lane", "value": gps,
        "unit": "grads/s", "kernel_ms": k_ms, "end_to_end_ms_incl_tape_build": wall * 1e3, "end_to_end_grads_per_s": lanes / wall,
        "rerun_ms_wall": rerun_wall * 1e3, "rerun_grads_per_s_wall": lanes / rerun_wall,
        "program_instructions": stats["instructions"], "program_slots": stats["slots"],
        "roofline": {"bound": "hbm", "achieved": gps * alg / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": gps * alg / 1e9 / pk["hbm_gbs"], "traffic": ncu_traffic("lane_program_c2"),
                     "algorithmic_bytes_per_gradient": alg, "algorithmic_bytes": alg * lanes}}
    del cols, res

    out["hvp"] = hvp_metric(ctx, ext, torch, wl, None, 1, 0)
    try:  # launch-bound sizes: the same step eager and as ONE replayed CUDA graph (gadcu_capture_*, north-star subsystem 1)
        out["graph_replay"] = replay_metric(ctx, ext, torch, gb, wl)
    except Exception as e:
        out["graph_replay"] = {"error": repr(e)}
    try:  # the CPU port beside each number (rank 0, one GPU only: this function runs there)
        cpu = cpu_secondary(host_cores())
        out["grad_sweep"]["cpu_baseline"] = cpu["c3"]
        out["batched_tape"]["cpu_baseline"] = cpu["c2"]
        out["scalar_tape"] = {"metric": "scalar-tape grads/s", "workload": "C1 one Rosenbrock-64 f64 Graph1 tape at a time (CPU only: "
                              "the device result for a single tape is a parity check, tests/test_gpu_workloads.py, not a throughput claim)",
                              "cpu_baseline": cpu["c1"]}
    except Exception as e:  # (never lose the bench line over a baseline)
        out["cpu_baseline_error"] = repr(e)
    return out


def replay_metric(ctx, ext, torch, gb, wl):
    """Where launch overhead, not the GPU, bounds a step: a small C4-shaped MLP step (B=512, D=256: 29 launches of a few
    microseconds each) and the C3 tape at n = 2^16, issued eagerly through the API and replayed as one captured graph.
    Device time per step with CUDA events on the context's stream, 200 steps after warm-up."""
    def ev_time(fn, reps):
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(reps):
            fn()
        e1.record(ext)
        e1.synchronize()
        return e0.elapsed_time(e1) / reps

    res = {}
    B_, D_ = 512, 256
    X = ctx.random((B_, D_), 5, normal=True, a=0.0, b=1.0)
    T = ctx.random((B_, D_), 9, normal=True, a=0.0, b=1.0)
    Ws = [ctx.random((D_, D_), 6 + l, normal=True, a=0.0, b=1.0 / np.sqrt(D_)) for l in range(L)]
    bs = [ctx.array(np.full((1, D_), 0.01, np.float32)) for _ in range(L)]
    keep = {}

    def eager():
        keep["e"] = wl.mlp_step(ctx, X, Ws, bs, T)

    for _ in range(5):
        eager()
    ms_eager = ev_time(eager, 200)
    with gb.Capture(ctx) as cap:
        captured = wl.mlp_step(ctx, X, Ws, bs, T)
    for _ in range(5):
        cap.launch()
    ms_replay = ev_time(cap.launch, 200)
    same = captured[0].item() == keep["e"][0].item()
    res["mlp_small"] = {"workload": f"C4-shaped step at B={B_}, D={D_}, L={L}, f32", "eager_ms": ms_eager, "replay_ms": ms_replay,
                        "eager_steps_per_s": 1e3 / ms_eager, "replay_steps_per_s": 1e3 / ms_replay, "loss_identical": bool(same)}
    del cap, captured
    n = 1 << 16
    x = ctx.random((n,), 3, a=-1.0, b=1.0)
    y = ctx.random((n,), 4, a=0.5, b=2.0)

    def eager3():
        keep["c3"] = wl.c3_step(ctx, x, y)

    for _ in range(5):
        eager3()
    ms_eager3 = ev_time(eager3, 200)
    with gb.Capture(ctx) as cap3:
        captured3 = wl.c3_step(ctx, x, y)
    for _ in range(5):
        cap3.launch()
    ms_replay3 = ev_time(cap3.launch, 200)
    res["c3_small"] = {"workload": f"C3 tape at n={n}, f32, forward + reverse sweep", "eager_ms": ms_eager3, "replay_ms": ms_replay3,
                       "z_identical": bool(captured3[0].item() == keep["c3"][0].item())}
    return res


def hvp_metric(ctx, ext, torch, wl, comm, world, rank, dist=None):
    """C5: Hessian-vector products/s of the MLP loss on GraphN (reverse over reverse), C4 shapes, rows sharded over
    the ranks and the products all-reduced. Every rank runs it; rank 0 reports."""
    try:
        rows = B_TOTAL // world
        X = ctx.random((rows, D), 5 + 100 * rank, normal=True, a=0.0, b=1.0)
        T = ctx.random((rows, D), 9 + 100 * rank, normal=True, a=0.0, b=1.0)
        Ws = [ctx.random((D, D), 6 + l, normal=True, a=0.0, b=1.0 / np.sqrt(D)) for l in range(L)]
        bs = [ctx.array(np.full((1, D), 0.01, np.float32)) for _ in range(L)]
        vs = [ctx.random((D, D), 10 + l, normal=True, a=0.0, b=1.0) for l in range(L)]
        hv = {}

        def hvp():
            hv["r"] = wl.mlp_hvp(ctx, X, Ws, bs, T, vs, comm)

        for _ in range(2):
            hvp()
        ctx.synchronize()
        if dist is not None:
            dist.barrier()
        ctx.profile_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(5):
            hvp()
        e1.record(ext)
        e1.synchronize()
        ms_hvp = e0.elapsed_time(e1) / 5
        prof = ctx.profile_end()
        if dist is not None:
            t = torch.tensor([ms_hvp], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_hvp = float(t.item())
        gemms = prof["matmul"]["launches"] / 5
        return {"metric": "MLP loss Hessian-vector products/s", "workload": f"C5 GraphN reverse-over-reverse, B={B_TOTAL}, D={D}, L={L}, f32",
                "value": 1000.0 / ms_hvp, "unit": "HVPs/s", "n_gpus": world, "ms_per_hvp": ms_hvp, "tape_nodes": int(hv["r"][2]),
                "gemms_per_hvp": gemms, "logical_tflop_per_hvp": gemms * 2.0 * B_TOTAL * D * D / 1e12,
                "gemm_ms_per_hvp_per_gpu": prof["matmul"]["ms"] / 5}
    except Exception as e:
        return {"error": repr(e)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--headline-only", action="store_true", help="skip the secondary metrics and the CPU baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
