"""CPU: the parts of bench.py's contract that need no GPU — the reference arm (`--impl reference`: the CPU port on the host
cores, one JSON line, rank 0 alone under torchrun) and the refusal of our arm to run without a device (no CPU fallback)."""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ["impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "cpu_baseline", "e2e"]


def _json_lines(stdout):
    out = []
    for line in stdout.splitlines():
        line = line.strip()
        if line.startswith("{"):
            out.append(json.loads(line))
    return out


def _check_reference_line(line, n_gpus):
    for k in REQUIRED:
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "MLP fwd+bwd steps/s" and line["unit"] == "steps/s"
    assert line["n_gpus"] == n_gpus and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["value"] > 0 and abs(line["ms_per_step"] * line["value"] - 1000.0) < 1e-6 * 1000.0
    assert line["config"]["workload"].startswith("C4 ") and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "batch rows" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_prints_one_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = _json_lines(res.stdout)
    assert len(lines) == 1
    _check_reference_line(lines[0], 1)


def test_reference_arm_under_torchrun_only_rank0_works():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", "2", "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = _json_lines(res.stdout)
    assert len(lines) == 1, res.stdout
    _check_reference_line(lines[0], 2)


def test_our_arm_refuses_to_run_without_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            return
    except Exception:
        pass
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert res.returncode != 0
    assert not _json_lines(res.stdout), "a bench line without a GPU would be a CPU fallback"


def test_committed_bench_lines_carry_the_whole_contract():
    """profiles/r01_bench_*.json are the lines the B200 runs printed: every key the contract names is there, with the
    roofline / cpu_baseline / e2e / clocks objects filled and no model keys in `config`."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r01_bench_final.json")) + glob.glob(os.path.join(ROOT, "profiles", "r01_bench_n[248].json")))
    assert len(files) == 4
    for path in files:
        line = _json_lines(open(path).read())[-1]
        for k in ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                  "config", "clocks", "e2e", "gpu_launches", "roofline"]:
            assert k in line, (path, k)
        assert line["metric"] == "MLP fwd+bwd steps/s" and line["unit"] == "steps/s" and line["scaling"] == "strong" and line["dtype"] == "f32"
        assert line["vs_baseline"] is None and line["data"] == "synthetic" and line["warmup"] >= 3 and line["gpu_launches"] > 0
        assert "workload" in line["config"] and "model" not in line["config"] and "l2" in line["config"]
        n = line["n_gpus"]
        assert os.path.basename(path) == ("r01_bench_final.json" if n == 1 else f"r01_bench_n{n}.json")
        r = line["roofline"]
        assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
        e = line["e2e"]
        assert e["unit"] == "steps/s" and 0 < e["value"] <= line["value"] * 1.02 and e["h2d_bytes_per_step"] == 8192 * 4096 * 4 * 2 and e["d2h_bytes_per_step"] == 4 * n
        c = line["clocks"]
        assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"] and not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"]))
        if n == 1:
            cb = line["cpu_baseline"]
            assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and "sample" in cb
            sec = line["secondary"]
            for key in ["grad_sweep", "batched_tape", "hvp", "graph_replay", "scalar_tape"]:
                assert key in sec, key
            assert sec["grad_sweep"]["roofline"]["bound"] == "hbm" and sec["batched_tape"]["roofline"]["bound"] == "hbm"
            assert sec["grad_sweep"]["cpu_baseline"]["kind"] == "port" and sec["batched_tape"]["cpu_baseline"]["cores"] >= 1


def test_committed_round2_bench_lines_carry_parity_dp_check_and_the_secondary_rows():
    """profiles/r02_bench_n{1,2,4,8}.json: what VERDICT r01 asked of every line — one fixed roofline denominator with the
    source of `traffic`, a parity object per row, `dp_check` in every N > 1 line, C2 and C5 over the ranks."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_n[1248].json")))
    assert len(files) == 4
    for path in files:
        line = _json_lines(open(path).read())[-1]
        n = line["n_gpus"]
        assert os.path.basename(path) == f"r02_bench_n{n}.json"
        for k in ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                  "config", "clocks", "e2e", "gpu_launches", "roofline"]:
            assert k in line, (path, k)
        assert line["metric"] == "MLP fwd+bwd steps/s" and line["scaling"] == "strong" and line["vs_baseline"] is None and line["warmup"] >= 3
        r = line["roofline"]
        assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and "burst" in r["peak_source"]
        assert "frac_of_sustained_peak" in r and "frac_of_burst_peak" not in r  # one denominator, the other a side field
        e = line["e2e"]
        assert e["h2d_bytes_per_step"] == 8192 * 4096 * 4 * 2 and e["d2h_bytes_per_step"] == 4 * n and 0 < e["value"] <= line["value"] * 1.02
        assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(line["clocks"]["reasons"]))
        sec = line["secondary"]
        if n == 1:
            assert r["traffic"] > 0 and "@" in r["traffic_source"] and r["traffic_source"].startswith("profiles/")
            cb = line["cpu_baseline"]
            assert cb["kind"] == "port" and "sgemm" in cb["gemm"] and cb["sample"].startswith("8192 of 8192")
            assert line["parity"]["vs"] == "oracle" and 0 < line["parity"]["max_rel"] < 1e-3
            for key in ["grad_sweep", "batched_tape", "hvp", "graph_replay", "scalar_tape"]:
                assert key in sec, key
            assert sec["grad_sweep"]["parity"]["vs"] == "oracle" and sec["batched_tape"]["parity"]["bit_identical"] is True
            h = sec["hvp"]
            assert h["roofline"]["bound"] == "tensor" and h["cpu_baseline"]["kind"] == "port" and h["parity"]["max_rel"] < 1e-3
            bt = sec["batched_tape"]
            assert bt["compiled_host"]["same_bits_as_python_recorded"] is True and bt["compiled_host"]["ms_per_call"] < 1.0
            gr = sec["graph_replay"]["mlp_small"]
            assert gr["auto_loss_identical"] is True and gr["auto_replay_segments"]["instantiated"] == 0
        else:
            dc = line["dp_check"]
            assert dc["rows_per_gpu"] == 8192 // n and dc["fused_repeatable"] is True
            assert dc["fused_vs_nccl_max_rel"] < 1e-4 and dc["vs_full_batch_on_one_gpu"]["max_rel"] < 1e-4
            assert sec["hvp"]["n_gpus"] == n and sec["hvp"]["value"] > 0
            assert sec["batched_tape"]["n_gpus"] == n and sec["batched_tape"]["lanes_per_gpu"] == (1 << 20) // n
