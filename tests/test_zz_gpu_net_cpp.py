"""GPU: gad/tests/net.rs replayed through the C++ mirror of gad's network layer (include/gadcu_net.hpp -> gadcu.hpp -> C ABI
-> device backend): a hand-written Net and input.using(weight).map(matmul_nn) trained to the identity with
apply_gradient_step, weights through the reference's checkpoint bytes, evaluate / check, one two-example gradient step
against its closed form, copy-on-write weight snapshots. The program is examples/net_cpp.cc; it exits 0 when every
expectation holds. (The same header is checked over the CPU oracle's tape in tests/test_net_cpp.py; this file sorts last
so the kernels' parity tests are not held up behind a host-language mirror.)"""
import os
import shutil
import subprocess
import tempfile

import pytest

from test_net_cpp import build_device_example

pytestmark = pytest.mark.gpu


def test_reference_net_tests_through_the_cpp_mirror():
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "net_cpp")
        res = build_device_example(exe)
        assert res.returncode == 0, res.stderr
        run = subprocess.run([exe], capture_output=True, text=True, timeout=600)
        assert run.returncode == 0, run.stdout + run.stderr
        assert "all net tests passed" in run.stdout


def test_generic_program_on_scalar_and_batched_tapes_through_the_cpp_mirror():
    """examples/rosenbrock_cpp.cc: Rosenbrock-N as ONE generic program on gadcu::Eval, Graph1 (one-element
    arrays: the scalar tape) and BatchedGraph1 (one tape per lane): known answers, the f64 closed form over 4096 lanes,
    sampled lanes bit for bit against the scalar tape, Store::rerun == a freshly recorded tape."""
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "rosenbrock_cpp")
        res = build_device_example(exe, "rosenbrock_cpp.cc")
        assert res.returncode == 0, res.stderr
        run = subprocess.run([exe], capture_output=True, text=True, timeout=600)
        assert run.returncode == 0 and "batched tapes and closed form agree" in run.stdout, run.stdout + run.stderr


def test_data_parallel_net_steps_through_the_cpp_mirror():
    """gadcu::Comm + apply_gradient_step(net, ctx, lambda, batch, comm): one process per GPU (two ranks), the examples split
    unevenly (5 over 2), a rank without examples (1 over 2) — losses and weights against the same steps taken alone."""
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs on the box")
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "net_cpp")
        res = build_device_example(exe)
        assert res.returncode == 0, res.stderr
        id_file = os.path.join(tmp, "comm_id")
        procs = [subprocess.Popen([exe, "dp", str(r), "2", id_file], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
        outs = []
        for p in procs:
            try:
                outs.append(p.communicate(timeout=300)[0])
            except subprocess.TimeoutExpired:
                for q in procs:
                    q.kill()
                raise
        for r, (p, out) in enumerate(zip(procs, outs)):
            assert p.returncode == 0 and f"rank {r}: data-parallel net steps ok" in out, outs
