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
def oracle_module():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py
    return oracle_py


def cpu_gemm_kind(oracle_py, threads: int) -> str:
    """The reference's arrays are ArrayFire's, whose CPU backend hands matmul to a BLAS sgemm; the port does the same when
    the host has one (the OpenBLAS numpy bundles), else it falls back to its own blocked loop and says so."""
    path = oracle_py.find_blas()
    if path and oracle_py.set_blas(path, threads):
        return "sgemm of " + os.path.basename(path).split("-")[0] + " (OpenBLAS bundled with numpy), as ArrayFire-CPU calls a BLAS"
    return "the port's own column-blocked loop, NOT a BLAS (none found on this host): understates an ArrayFire-CPU build"


def synthetic_mlp_host(rows: int, seed_shift: int = 0):
    rng = np.random.default_rng(5 + seed_shift)
    X = rng.standard_normal((rows, D)).astype(np.float32)
    Ws = [(rng.standard_normal((D, D)) / np.sqrt(D)).astype(np.float32) for _ in range(L)]
    bs = [np.full((1, D), 0.01, np.float32) for _ in range(L)]
    T = rng.standard_normal((rows, D)).astype(np.float32)
    return X, Ws, bs, T


def cpu_mlp_step_seconds(oracle_py, rows: int, threads: int, data=None):
    X, Ws, bs, T = data if data is not None else synthetic_mlp_host(rows)
    loss, gW, gb_, secs = oracle_py.mlp_step(X, Ws, bs, T, nthreads=threads)
    return secs, (loss, gW, gb_)


def cpu_secondary(threads: int):
    """The restated CPU path beside the secondary metrics, each on a bounded sample (about a second of CPU work):
    C1 one scalar Rosenbrock-64 tape on one core, C2 the same tape looped over lanes on all cores, C3 the HostArray tape
    (one core: the reference's CPU arrays are single-threaded per op) with the reverse sweep timed on its own."""
    oracle_py = oracle_module()
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
    """--impl reference: the reference's CPU path for the same config and metric. The reference itself (Rust + the arrayfire
    crate) cannot be built in this image, so this is oracle/ (kind "port") with all host cores. Every timed step is one
    full C4 step over `rows` batch rows at full width; rows = the whole batch when the run then still ends within a few
    minutes, otherwise the largest power-of-two share that does (steps/s scaled by 8192 / rows: work is linear in rows)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    oracle_py = oracle_module()
    threads = host_cores()
    gemm = cpu_gemm_kind(oracle_py, threads)
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    t_cal, _ = cpu_mlp_step_seconds(oracle_py, 1024, threads)  # calibration (untimed): one eighth of the batch
    rows = B_TOTAL
    budget = float(os.environ.get("GADCU_REF_BUDGET_S", "200"))
    while rows > 1024 and (steps + warmup) * t_cal * (rows / 1024.0) > budget:
        rows //= 2
    data = synthetic_mlp_host(rows)
    for _ in range(warmup):
        cpu_mlp_step_seconds(oracle_py, rows, threads, data)
    times = [cpu_mlp_step_seconds(oracle_py, rows, threads, data)[0] for _ in range(steps)]
    scale = B_TOTAL // rows
    sec_per_step = float(np.mean(times)) * scale
    v = 1.0 / sec_per_step
    sample = (f"{rows} of {B_TOTAL} batch rows at full width D={D}, L={L}" +
              ("" if scale == 1 else f"; steps/s scaled by {scale}x (work is linear in rows)") + f"; matmul = {gemm}")
    line = {"impl": "reference", "metric": "MLP fwd+bwd steps/s", "value": v, "unit": "steps/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C4 3-layer tanh MLP {D}-wide, batch {B_TOTAL}, f32, Graph1 fwd + reverse sweep", "global_batch": B_TOTAL,
                       "rows_timed_per_step": rows},
            "cpu_baseline": {"value": v, "unit": "steps/s", "cores": threads, "kind": "port", "sample": sample, "gemm": gemm},
            "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference (Rust crate gad + arrayfire) cannot be built in this image; this is oracle/, a C++ port of its CPU tape"}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def ncu_traffic_entry(kernel: str):
    """(bytes per launch, "profiles/<capture>@<commit>") of `kernel` from the committed ncu --set full capture: the DRAM
    bytes are a property of the kernel at that commit, measured once under the profiler, not of this run."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            e = json.load(f).get(kernel)
        if isinstance(e, dict):
            return e.get("bytes"), f"profiles/{e.get('source')}@{e.get('commit')}"
        return e, "profiles/ncu_traffic.json"
    except Exception:
        return None, None


def rel_err(got, want) -> float:
    """max |got - want| / max |want| in f64."""
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    return float(np.abs(got - want).max() / max(float(np.abs(want).max()), 1e-300))


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

    # synthetic data of the named shapes: X~N(0,1) seed 5, W_l~N(0,1/D) seeds 6-8, b=0.01, target~N(0,1) seed 9;
    # rank r holds rows [r rows, (r+1) rows) of the batch, generated from its own seeds
    def shard(r):
        return (ctx.random((rows, D), 5 + 100 * r, normal=True, a=0.0, b=1.0), ctx.random((rows, D), 9 + 100 * r, normal=True, a=0.0, b=1.0))

    X, T = shard(rank)
    Ws = [ctx.random((D, D), 6 + l, normal=True, a=0.0, b=1.0 / np.sqrt(D)) for l in range(L)]
    bs = [ctx.array(np.full((1, D), 0.01, np.float32)) for _ in range(L)]

    def barrier():
        ctx.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = [0.0]

    def timed(fn, steps, drain=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        t_host = time.perf_counter()
        for _ in range(steps):
            fn()
        if drain is not None:
            drain()  # (what the loop still owes — the last step's read-back — belongs to the timed region)
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
    # brings step i+1's X / target rows from pinned host memory (the copy is ordered behind step i-1's reads of the
    # buffer it overwrites: gadcu_array_upload_async). Every timed step issues exactly one such pair of copies and one
    # device -> host copy of its own loss into pinned memory; the host picks the value up one step later
    # (gadcu_array_download_async / gadcu_ctx_wait_download), as a training loop logs its loss, so that it can issue
    # step i+1 while step i runs. The last step's loss is waited for inside the timed region (`drain`).
    # GADCU_E2E_SYNC_LOSS=1: the earlier loop (item() every step: the host waits for each step before issuing the next).
    bufs = [(ctx.empty((rows, D)), ctx.empty((rows, D))) for _ in range(2)]
    hloss = [torch.zeros((1,), dtype=torch.float32).pin_memory() for _ in range(2)]
    nloss = [h.numpy() for h in hloss]
    sync_loss = os.environ.get("GADCU_E2E_SYNC_LOSS", "0") == "1"
    e2e_loss, e2e_i, e2e_pending = {"n": 0}, [0], [None]

    def prefetch(i):
        bx, bt = bufs[i % 2]
        bx.upload_async(nX)
        bt.upload_async(nT)

    def collect():
        if e2e_pending[0] is not None:
            ticket, slot = e2e_pending[0]
            ctx.wait_download(ticket)
            e2e_loss["v"] = float(nloss[slot][0])
            e2e_loss["n"] += 1
            e2e_pending[0] = None

    def step_e2e():
        i = e2e_i[0]
        ctx.wait_uploads()       # this step's inputs (copied during the previous step)
        prefetch(i + 1)          # next step's inputs, overlapping this step's kernels
        bx, bt = bufs[i % 2]
        out = wl.mlp_step(ctx, bx, Ws, bs, bt, comm, overlap)
        if sync_loss:
            e2e_loss["v"] = out[0].item()  # device -> host read of the step's result, host waits
            e2e_loss["n"] += 1
        else:
            ticket = out[0].download_async(nloss[i % 2])  # queued behind this step's kernels
            collect()                                     # the previous step's loss: on the host by now
            e2e_pending[0] = (ticket, i % 2)
        e2e_i[0] = i + 1

    prefetch(0)
    for _ in range(3):
        step_e2e()
    collect()
    e2e_steps = max(1, args.steps)
    reads0 = e2e_loss["n"]
    ms_e2e = timed(step_e2e, e2e_steps, drain=collect)
    assert e2e_loss["n"] - reads0 == e2e_steps, "every timed step's loss must have reached the host inside the timed region"
    e2e = {"value": e2e_steps / (ms_e2e / 1000.0), "unit": "steps/s", "h2d_bytes_per_step": int(2 * rows * D * 4 * world),
           "d2h_bytes_per_step": int(4 * world), "ms_per_step": ms_e2e / e2e_steps, "loss": e2e_loss.get("v"),
           "loss_read": "item() every step" if sync_loss else "async copy queued by the step, value taken one step later; all inside the timed region"}
    del bufs

    pk = peaks()
    secondary = {}
    dp_check = None
    if world > 1:
        try:  # untimed: the exchange's result against NCCL and against the full batch on one GPU (every rank takes part)
            dp_check = dp_check_block(ctx, wl, comm, dist, world, rank, X, T, Ws, bs, shard, overlap)
        except Exception as e:
            dp_check = {"error": repr(e)}
    if not args.headline_only:
        if world > 1:
            secondary["hvp"] = hvp_metric(ctx, ext, torch, wl, comm, world, rank, pk, dist)
            try:
                secondary["batched_tape"] = batched_metric(ctx, torch, gb, wl, pk, world, rank, dist)[0]
            except Exception as e:
                secondary["batched_tape"] = {"error": repr(e)}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # TF32 dense peak = half of the measured bf16 figure. The timed region is a fraction of a second of GEMMs, so the
    # denominator is the BURST figure of MEASURED_PEAKS.json, always; the sustained one (seconds under the power cap) is a
    # side field. (r01 switched between the two on the sampled clock, which gave one kernel two fractions.)
    tf32_peak = pk["bf16_tflops"] / 2.0
    gemm = prof["matmul"]
    gemm_ms = gemm["ms"] / max(1, gemm["launches"])
    logical_flop = 2.0 * rows * D * D
    mma_flop = 3.0 * logical_flop  # 3xTF32: three tensor-core products per logical product
    achieved = mma_flop / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    traffic, traffic_source = ncu_traffic_entry("gemm_tf32x3_pair_kernel")
    roofline = {"bound": "tensor", "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved / tf32_peak,
                "traffic": traffic if world == 1 else None, "traffic_source": traffic_source if world == 1 else None,
                "kernel": "gemm_tf32x3_pair_kernel (fwd / dA / dW) incl. whatever operand preparation runs as separate launches, 3xTF32",
                "frac_of_sustained_peak": achieved / (pk["bf16_tflops_sustained"] / 2.0),
                "logical_fp32_tflops": logical_flop / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0,
                "peak_source": f"{pk['source']}: bf16_tflops/2 (burst figure; TF32 dense = half of bf16 dense)", "avg_launch_ms": gemm_ms,
                "launches": gemm["launches"], "flop_per_launch": mma_flop, "share_of_step": gemm["ms"] / ms if ms > 0 else None}

    line = {"metric": "MLP fwd+bwd steps/s", "value": steps_per_s, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C4 3-layer tanh MLP {D}-wide, batch {B_TOTAL}, f32, Graph1 fwd + reverse sweep", "global_batch": B_TOTAL,
                       "rows_per_gpu": rows, "parallelism": f"dp{world} (batch rows sharded, weight gradients summed over ranks" + (": dW GEMMs feed a reduce-scatter / all-gather over NVLink peer memory overlapped with the sweep)" if overlap and world > 1 else ")"),
                       "l2": "inputs larger than L2 (activations 128 MiB, weights 64 MiB each vs 126 MB L2)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "kernel_time_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()}, "host_issue_ms_per_step": host_issue_ms,
            "loss": loss_val}
    if dp_check is not None:
        line["dp_check"] = dp_check

    if world == 1 and not args.headline_only:
        try:
            secondary = secondary_metrics(ctx, ext, torch, gb, wl, pk, args)
        except Exception as e:  # the headline line must survive
            secondary = {"error": repr(e)}
        try:
            line["cpu_baseline"], line["parity"] = cpu_baseline_and_parity(ctx, wl, X, T, Ws, bs, keep["out"])
        except Exception as e:
            line["cpu_baseline"] = {"error": repr(e)}
    if secondary:
        line["secondary"] = secondary
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def cpu_baseline_and_parity(ctx, wl, X, T, Ws, bs, device_out):
    """The CPU port on the SAME inputs as the device step (read back from the device), all host cores, at the full batch
    when that is about half a minute of CPU work or less, else on the first `rows` rows (then the device recomputes that
    shard for the comparison). One run gives both the reported baseline and the parity error of this line."""
    oracle_py = oracle_module()
    threads = host_cores()
    gemm = cpu_gemm_kind(oracle_py, threads)
    hW = [w.numpy()[:, :, 0, 0] for w in Ws]
    hb = [b.numpy()[:, :, 0, 0] for b in bs]
    hX, hT = X.numpy()[:, :, 0, 0], T.numpy()[:, :, 0, 0]
    t_cal, _ = cpu_mlp_step_seconds(oracle_py, 512, threads, (np.ascontiguousarray(hX[:512]), hW, hb, np.ascontiguousarray(hT[:512])))
    rows = B_TOTAL
    while rows > 1024 and t_cal * (rows / 512.0) > 30.0:
        rows //= 2
    data = (np.ascontiguousarray(hX[:rows]), hW, hb, np.ascontiguousarray(hT[:rows]))
    secs, (oloss, ogW, ogb) = cpu_mlp_step_seconds(oracle_py, rows, threads, data)
    scale = B_TOTAL // rows
    if rows == B_TOTAL:
        dloss, dgW, dgb = device_out
    else:
        dloss, dgW, dgb = wl.mlp_step(ctx, ctx.array(data[0]), Ws, bs, ctx.array(data[3]))
    per_array = {f"dW{l + 1}": rel_err(dgW[l].numpy()[:, :, 0, 0], ogW[l]) for l in range(L)}
    per_array.update({f"db{l + 1}": rel_err(dgb[l].numpy()[:, :, 0, 0], ogb[l]) for l in range(L)})
    parity = {"max_rel": max(per_array.values()), "vs": "oracle", "per_array": per_array,
              "loss": {"device": dloss.item(), "oracle": oloss,
                       "note": "the oracle adds the 33.5 M squared residuals one after the other in f32 (its documented order), which loses whole terms once the "
                               "running sum passes 2^24; the device's tree sum is the one that agrees with f64 (tests/test_gpu_parity_full.py: 2e-5)"},
              "rows": rows, "what": "max |device - oracle| / max |oracle| over every dL/dW_l and dL/db_l of one step on the same inputs; both are f32 "
              "computations with different summation orders (3xTF32 tensor-core tiles vs a CPU sgemm), so this is f32 rounding of sums over "
              f"{rows} rows, not a bit comparison"}
    cpu = {"value": 1.0 / (secs * scale), "unit": "steps/s", "cores": threads, "kind": "port", "seconds": secs, "gemm": gemm,
           "sample": f"{rows} of {B_TOTAL} batch rows at full width D={D}, L={L}, the device's own inputs" + ("" if scale == 1 else f"; scaled by {scale}x")}
    return cpu, parity


def dp_check_block(ctx, wl, comm, dist, world, rank, X, T, Ws, bs, shard, overlap):
    """Untimed correctness of the N-GPU step at the bench's own shard shape (rows = 8192 / N, D = 4096):
      * the fused exchange against one grouped NCCL all-reduce of the same local gradients (bit-equal at 2 ranks when no GEMM
        split its k range: a + b is the same in either order; beyond that the association differs);
      * against the full batch evaluated on rank 0 alone, shard by shard with the gradients added up on that one GPU."""
    a = wl.mlp_step(ctx, X, Ws, bs, T, comm, overlap)
    b = wl.mlp_step(ctx, X, Ws, bs, T, comm, False)
    a_np = [g.numpy() for g in a[1] + a[2]]
    b_np = [g.numpy() for g in b[1] + b[2]]
    names = [f"dW{l + 1}" for l in range(L)] + [f"db{l + 1}" for l in range(L)]
    out = {"rows_per_gpu": int(X.dims()[0]), "fused_vs_nccl_max_rel": max(rel_err(x, y) for x, y in zip(a_np, b_np)),
           "fused_vs_nccl_bit_equal": bool(all(np.array_equal(x, y) for x, y in zip(a_np, b_np)) and a[0].item() == b[0].item())}
    c = wl.mlp_step(ctx, X, Ws, bs, T, comm, overlap)
    out["fused_repeatable"] = bool(all(np.array_equal(x, g.numpy()) for x, g in zip(a_np, c[1] + c[2])))
    if rank == 0:
        tot_loss, acc = 0.0, None
        for r in range(world):
            Xr, Tr = (X, T) if r == 0 else shard(r)
            l_, gW, gb_ = wl.mlp_step(ctx, Xr, Ws, bs, Tr)
            tot_loss += l_.item()
            if acc is None:
                acc = [g.clone() for g in gW + gb_]
            else:
                for dst, src in zip(acc, gW + gb_):
                    dst.add_assign(src)
        per = {n: rel_err(x, g.numpy()) for n, x, g in zip(names, a_np, acc)}
        out["vs_full_batch_on_one_gpu"] = {"max_rel": max(per.values()), "per_array": per, "loss_rel": abs(a[0].item() - tot_loss) / abs(tot_loss)}
    ctx.synchronize()
    dist.barrier()
    return out


def batched_metric(ctx, torch, gb, wl, pk, world, rank, dist):
    """C2: batched scalar tapes, 2^20 lanes x Rosenbrock-64, f64, one tape per lane; with N ranks the lanes are split evenly
    and nothing is exchanged (independent units). `value` is what a gad program gets through the API: record the tape
    (569 nodes), sweep, launch — every call, as the reference re-records its tape per example (net_ext.rs:52)."""
    lanes_total, N = 1 << 20, 64
    lanes = lanes_total // world
    cols = [ctx.random((lanes,), 200 + i + 1000 * rank, dtype=gb.F64, a=-1.2, b=1.2) for i in range(N)]
    res = {}

    def batched():
        res["r"] = wl.rosenbrock_batched_gradients(ctx, cols)

    for _ in range(3):
        batched()
    ctx.synchronize()
    if dist is not None:
        dist.barrier()
    ctx.profile_begin()
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        batched()
    ctx.synchronize()
    wall = (time.perf_counter() - t0) / reps
    prof = ctx.profile_end()
    k_ms = prof["tile_transpose"]["ms"] / max(1, prof["tile_transpose"]["launches"])
    if dist is not None:
        t = torch.tensor([wall, k_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall, k_ms = float(t[0].item()), float(t[1].item())
    stats = res["r"][3].program_stats()
    alg = 2 * N * 8  # bytes per gradient: read 64 f64, write 64 f64
    kernel_gps = lanes_total / (k_ms * 1e-3) if k_ms > 0 else 0.0
    traffic, traffic_source = ncu_traffic_entry("lane_program_c2")
    # the same call sequence from a COMPILED host (examples/c2_host.c over include/gadcu.h: what a Rust gad program pays per step,
    # without ctypes' microseconds per call), on this rank's GPU and this rank's columns; the slowest rank counts
    compiled = None
    exe = os.path.join(ROOT, "examples", "c2_host")
    if os.path.exists(exe):
        try:
            ctx.synchronize()
            r = subprocess.run([exe, str(ctx.device), str(lanes), "20", str(200 + 1000 * rank)], capture_output=True, text=True, timeout=300)
            compiled = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": r.stderr[-500:]}
        except Exception as e:
            compiled = {"error": repr(e)}
        if compiled and "ms_per_call" in compiled:
            grads = res["r"][2]
            bits = 0
            for c in grads:
                bits = (bits + int(c.numpy()[:4096, 0, 0, 0].view(np.uint64).sum(dtype=np.uint64))) % (1 << 64)
            compiled["same_bits_as_python_recorded"] = bool(bits == int(compiled.pop("bits_checksum_first_4096_lanes")))
            c_ms = compiled["ms_per_call"]
            if dist is not None:
                t = torch.tensor([c_ms], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                c_ms = float(t[0].item())
            compiled["ms_per_call_max_over_ranks"] = c_ms
            compiled["grads_per_s"] = lanes_total / (c_ms * 1e-3)
    elif dist is not None:
        pass
    api_gps = compiled["grads_per_s"] if compiled and "grads_per_s" in compiled and "error" not in compiled else lanes_total / wall
    out = {"metric": "batched-tape grads/s", "workload": f"C2 {lanes_total} lanes x Rosenbrock-{N}, f64, one tape per lane", "n_gpus": world,
           "lanes_per_gpu": lanes, "value": api_gps, "unit": "grads/s",
           "value_is": "API level: every call records the 569-node tape, sweeps it and launches (the tape is re-recorded per step, as the reference does: "
                       "net_ext.rs:52)" + ("; host = compiled C over the C ABI (examples/c2_host.c)" if api_gps != lanes_total / wall else "; host = Python (ctypes)"),
           "compiled_host": compiled, "python_host": {"ms_per_call": wall * 1e3, "grads_per_s": lanes_total / wall,
                                                      "note": "the same calls through ctypes: ~1 100 calls x several microseconds of marshalling each"},
           "kernel_ms": k_ms, "kernel_grads_per_s": kernel_gps, "scaling": "total lanes fixed, split evenly over the ranks, no collective" if world > 1 else None,
           "program_instructions": stats["instructions"], "program_slots": stats["slots"],
           "roofline": {"bound": "hbm", "achieved": kernel_gps * alg / 1e9 / world, "peak": pk["hbm_gbs"], "unit": "GB/s",
                        "frac": kernel_gps * alg / 1e9 / world / pk["hbm_gbs"], "traffic": traffic if world == 1 else None,
                        "traffic_source": traffic_source if world == 1 else None, "algorithmic_bytes_per_gradient": alg,
                        "algorithmic_bytes": alg * lanes, "note": "per GPU: the lane kernel's algorithmic bytes / its CUDA-event time"}}
    return out, res, cols


def secondary_metrics(ctx, ext, torch, gb, wl, pk, args):
    out = {}
    oracle_py = oracle_module()

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
    traffic, traffic_source = ncu_traffic_entry("lane_program_c3_sweep")
    out["grad_sweep"] = {
        "metric": "grad-sweep HBM GB/s", "workload": "C3 64Mi-element exp/log/mul/add + sum, f32, reverse sweep", "value": gbs, "unit": "GB/s",
        "ms_per_sweep": ms_sweep, "launches_per_sweep": int(sweep_launches),
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                     "frac_of_8TBs": gbs / 8000.0, "traffic": traffic, "traffic_source": traffic_source, "algorithmic_bytes": alg_bytes,
                     "note": "algorithmic = 4 passes (read x,y; write gx,gy); the per-node sweep materialises more (see DESIGN.md)"},
        "fwd_plus_sweep_ms": ms_fb, "fwd_plus_sweep_gbs": 6 * n * 4 / (ms_fb * 1e-3) / 1e9}
    try:  # parity of this row: the oracle on a 2^22-element prefix of the same x, y (elementwise: prefix gradients do not depend on the rest)
        m = 1 << 22
        hx, hy = x.numpy().reshape(-1)[:m].copy(), y.numpy().reshape(-1)[:m].copy()
        _, ogx, ogy, _, _ = oracle_py.c3(hx, hy)
        dgx, dgy = keep["r"][1].numpy().reshape(-1)[:m], keep["r"][2].numpy().reshape(-1)[:m]
        eps = float(np.finfo(np.float32).eps)
        out["grad_sweep"]["parity"] = {"vs": "oracle", "elements": m,
                                       "max_rel_dz_dy": float(np.max(np.abs(dgy - ogy) / np.abs(ogy))) / eps,
                                       "max_err_dz_dx_in_eps_of_summands": float(np.max(np.abs(dgx.astype(np.float64) - ogx) / (np.abs(ogx.astype(np.float64) - 1.0) + 1.0))) / eps,
                                       "unit": "f32 eps", "what": "dz/dy = exp(x)/y relative; dz/dx = exp(x) log(y) + 1 relative to |exp(x) log(y)| + 1 (the sum cancels); CUDA expf/logf vs glibc"}
    except Exception as e:
        out["grad_sweep"]["parity"] = {"error": repr(e)}
    del keep, g, vx, vy, z, x, y

    # ---- C2
    bt, res, cols = batched_metric(ctx, torch, gb, wl, pk, 1, 0, None)
    lanes, N = 1 << 20, 64
    # steady state of an optimiser loop that keeps the store: the same tapes at new points through the compiled program
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
    bt["rerun_ms_wall"] = rerun_wall * 1e3
    bt["rerun_grads_per_s_wall"] = lanes / rerun_wall
    try:  # parity of this row: sampled lanes against the oracle's scalar tape, bit for bit
        idx = np.random.default_rng(0).choice(lanes, 1024, replace=False)
        hx = np.stack([c.numpy()[:, 0, 0, 0][idx] for c in cols])
        got = np.stack([c.numpy()[:, 0, 0, 0][idx] for c in res["r"][2]])
        _, og, _ = oracle_py.rosenbrock_batch(hx, nthreads=host_cores())
        bt["parity"] = {"vs": "oracle", "lanes_sampled": int(idx.size), "max_ulp": 0 if np.array_equal(got, og) else int(np.max(np.abs(got.view(np.int64) - og.view(np.int64)))),
                        "bit_identical": bool(np.array_equal(got, og))}
    except Exception as e:
        bt["parity"] = {"error": repr(e)}
    out["batched_tape"] = bt
    del cols, res, cols2, keep_r, store

    out["hvp"] = hvp_metric(ctx, ext, torch, wl, None, 1, 0, pk)
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
        ctx.flush()  # (auto replay: what the last call recorded reaches the stream before the closing event)
        e1.record(ext)
        e1.synchronize()
        return e0.elapsed_time(e1) / reps

    def auto_time(fn, reps):
        """The same plain calls with gadcu_ctx_set_auto_replay on: every step rebuilds its tape; what it issues runs as one
        graph laid over the last step's."""
        ctx.set_auto_replay(True)
        try:
            for _ in range(5):
                fn()
            s0 = ctx.replay_stats()
            ms = ev_time(fn, reps)
            s1 = ctx.replay_stats()
        finally:
            ctx.set_auto_replay(False)
        return ms, {k: s1[k] - s0[k] for k in s0}

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
    ms_auto, auto_stats = auto_time(eager, 200)
    res["mlp_small"] = {"workload": f"C4-shaped step at B={B_}, D={D_}, L={L}, f32", "eager_ms": ms_eager, "replay_ms": ms_replay,
                        "auto_replay_ms": ms_auto, "auto_replay_segments": auto_stats,
                        "eager_steps_per_s": 1e3 / ms_eager, "replay_steps_per_s": 1e3 / ms_replay, "auto_replay_steps_per_s": 1e3 / ms_auto,
                        "loss_identical": bool(same), "auto_loss_identical": bool(keep["e"][0].item() == captured[0].item()),
                        "note": "replay = an explicit gadcu_capture of one step launched again; auto_replay = plain wl.mlp_step calls (a fresh Graph1 each) "
                                "with gadcu_ctx_set_auto_replay: the recorded kernel sequence is the structural key"}
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
    z_eager = keep["c3"][0].item()
    ms_auto3, auto_stats3 = auto_time(eager3, 200)
    res["c3_small"] = {"workload": f"C3 tape at n={n}, f32, forward + reverse sweep", "eager_ms": ms_eager3, "replay_ms": ms_replay3,
                       "auto_replay_ms": ms_auto3, "auto_replay_segments": auto_stats3,
                       "z_identical": bool(captured3[0].item() == z_eager), "auto_z_identical": bool(keep["c3"][0].item() == z_eager)}
    return res


def hvp_metric(ctx, ext, torch, wl, comm, world, rank, pk, dist=None):
    """C5: Hessian-vector products/s of the MLP loss on GraphN (reverse over reverse), C4 shapes, rows sharded over
    the ranks and the products summed over them as the second sweep completes each. Every rank runs it; rank 0 reports.
    The reference's own matmul VJP (matrix.rs:164-176) cannot produce this workload for B != D — it is a Dimensions error
    for the transposed operands of the recorded backward GEMMs — so what runs here is the library's default (corrected
    adjoint) mode, and the CPU figure beside it is the oracle with the same switch."""
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
        reps = 5
        ctx.profile_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(reps):
            hvp()
        e1.record(ext)
        e1.synchronize()
        ms_hvp = e0.elapsed_time(e1) / reps
        prof = ctx.profile_end()
        if dist is not None:
            t = torch.tensor([ms_hvp], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_hvp = float(t.item())
        gemms = prof["matmul"]["launches"] / reps
        gemm_ms = prof["matmul"]["ms"] / max(1, prof["matmul"]["launches"])
        mma_flop = 3.0 * 2.0 * rows * D * D  # per GEMM launch: every product of both sweeps is [rows or D] x D x [D or rows]
        achieved = mma_flop / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        out = {"metric": "MLP loss Hessian-vector products/s", "workload": f"C5 GraphN reverse-over-reverse, B={B_TOTAL}, D={D}, L={L}, f32",
               "value": 1000.0 / ms_hvp, "unit": "HVPs/s", "n_gpus": world, "ms_per_hvp": ms_hvp, "tape_nodes": int(hv["r"][2]),
               "gemms_per_hvp": gemms, "logical_tflop_per_hvp": gemms * 2.0 * B_TOTAL * D * D / 1e12,
               "gemm_ms_per_hvp_per_gpu": prof["matmul"]["ms"] / reps,
               "kernel_time_ms_per_hvp": {k: v["ms"] / reps for k, v in prof.items()},
               "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops"] / 2.0, "unit": "TFLOP/s",
                            "frac": achieved / (pk["bf16_tflops"] / 2.0), "traffic": None, "avg_launch_ms": gemm_ms,
                            "flop_per_launch": mma_flop, "share_of_hvp": prof["matmul"]["ms"] / reps / ms_hvp,
                            "kernel": "gemm_tf32x3_pair_kernel, the same kernel as the headline (3xTF32 MMA flop / CUDA-event time of the GEMM launches)"},
               "note": "the reference's matmul VJP cannot differentiate through its own backward GEMMs for B != D (gad/src/matrix.rs:164-176: "
                       "Dimensions error for transposed operands); measured in the library's default corrected-adjoint mode"}
        if world == 1 and rank == 0:
            try:
                out["cpu_baseline"], out["parity"] = hvp_cpu_baseline_and_parity(ctx, wl, X, T, Ws, bs, vs)
            except Exception as e:
                out["cpu_baseline"] = {"error": repr(e)}
        return out
    except Exception as e:
        return {"error": repr(e)}


def hvp_cpu_baseline_and_parity(ctx, wl, X, T, Ws, bs, vs):
    """The oracle's GraphN with the corrected adjoint on the first 1024 rows of the device's own inputs (all host cores,
    sgemm through the host's BLAS when there is one), timed; the device recomputes that shard for the comparison."""
    oracle_py = oracle_module()
    threads = host_cores()
    gemm = cpu_gemm_kind(oracle_py, threads)
    rows = 1024
    hX, hT = np.ascontiguousarray(X.numpy()[:rows, :, 0, 0]), np.ascontiguousarray(T.numpy()[:rows, :, 0, 0])
    hW, hb, hv_ = [w.numpy()[:, :, 0, 0] for w in Ws], [b.numpy()[:, :, 0, 0] for b in bs], [v.numpy()[:, :, 0, 0] for v in vs]
    oloss, ohv, secs, nodes = oracle_py.mlp_hvp(hX, hW, hb, hT, hv_, nthreads=threads)
    dhv, dloss, dnodes = wl.mlp_hvp(ctx, ctx.array(hX), Ws, bs, ctx.array(hT), vs)
    per = {f"Hv{l + 1}": rel_err(dhv[l].numpy()[:, :, 0, 0], ohv[l]) for l in range(L)}
    scale = B_TOTAL // rows
    cpu = {"value": 1.0 / (secs * scale), "unit": "HVPs/s", "cores": threads, "kind": "port", "seconds": secs, "gemm": gemm,
           "sample": f"{rows} of {B_TOTAL} batch rows at full width, the device's own inputs; scaled by {scale}x; oracle GraphN with the corrected matmul adjoint"}
    parity = {"max_rel": max(per.values()), "vs": "oracle (corrected adjoint)", "per_array": per, "rows": rows,
              "tape_nodes_equal": bool(int(dnodes) == int(nodes))}
    return cpu, parity


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
