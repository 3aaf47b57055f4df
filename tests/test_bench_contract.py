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
