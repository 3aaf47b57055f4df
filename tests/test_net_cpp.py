"""CPU: the C++ mirror of gad's network layer (include/gadcu_net.hpp — net.rs / net_ext.rs: combinators, WeightOps over
weight structures, apply_gradient_step, the bincode checkpoint layout), SURVEY §8(f) row 1 in the reference's own kind of
language. The header is generic over the algebra like the reference, so tests/cpp/net_test.cc instantiates it over the CPU
oracle's tape and replays gad/tests/net.rs without a GPU; the device instantiation (examples/net_cpp.cc) is compiled here
warning-free and must fail loudly without a GPU; on a B200 it runs in tests/test_zz_gpu_net_cpp.py."""
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

from gad_b200 import _lib, net

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="no g++")


def build_device_example(out_path, source="net_cpp.cc"):
    return subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", os.path.join(ROOT, "examples", source),
                           "-I", os.path.join(ROOT, "include"), "-L", os.path.dirname(_lib.LIB_PATH), "-lgadcu",
                           "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH), "-o", out_path], capture_output=True, text=True)


@pytest.fixture(scope="module")
def net_test_exe():
    _lib.load()
    tmp = tempfile.mkdtemp()
    exe = os.path.join(tmp, "net_test")
    base = ["g++", "-std=c++20", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "oracle"),
            os.path.join(ROOT, "tests", "cpp", "net_test.cc"), "-L", os.path.dirname(_lib.LIB_PATH), "-lgadcu",
            "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH), "-o", exe]
    res = subprocess.run(base + ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"], capture_output=True, text=True)
    if res.returncode != 0 and "sanitize" in res.stderr:  # (no sanitizer runtime in the image)
        res = subprocess.run(base, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    yield exe
    shutil.rmtree(tmp, ignore_errors=True)


def test_reference_net_tests_over_the_oracle_tape(net_test_exe):
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0")  # (libgadcu.so's static CUDA runtime keeps process-lifetime allocations)
    run = subprocess.run([net_test_exe], capture_output=True, text=True, timeout=600, env=env)
    assert run.returncode == 0 and run.stdout.strip().endswith("ok (0 failures)"), run.stdout + run.stderr


def test_checkpoint_bytes_equal_the_python_mirror(net_test_exe):
    """One layout, two writers: gadcu::net::serialize_weights_bincode and gad_b200.net.serialize_weights_bincode produce
    the same bytes for `((), W)` (the weights of input.using(weight)) and for `([W, I], d)` — a Vec is length-prefixed,
    tuples and `()` add nothing, f32 is af::DType 0 and f64 is 2 — and the Python reader takes the C++ bytes back."""
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0")
    run = subprocess.run([net_test_exe, "hex"], capture_output=True, text=True, timeout=60, env=env)
    assert run.returncode == 0, run.stdout + run.stderr
    first, second = (bytes.fromhex(line) for line in run.stdout.split())
    w0 = np.array([0.3, -0.2, 0.5, 0.1, 0.4, -0.6, -0.7, 0.2, 0.9], np.float32).reshape((3, 3), order="F")
    eye = np.eye(3, dtype=np.float32)
    d = np.array([0.25, -3.0], np.float64).reshape((2, 1), order="F")
    assert net.serialize_weights_bincode(((), w0)) == first
    assert net.serialize_weights_bincode(([w0, eye], d)) == second
    back = net.deserialize_weights_bincode(second, ([w0, eye], d))
    assert np.array_equal(back[0][0][:, :, 0, 0], w0) and np.array_equal(back[0][1][:, :, 0, 0], eye) and np.array_equal(back[1][:, :, 0, 0], d)


def test_shares_of_a_batch_equal_the_python_mirror(net_test_exe):
    """gadcu::net::shard_bounds (who works on which examples of a data-parallel step) against gad_b200.dp.shard_bounds."""
    from gad_b200 import dp
    cases = [(t, w) for w in (1, 2, 3, 4, 7, 8) for t in (0, 1, 2, 5, 8, 63, 64, 1000, 8192)]
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0")
    run = subprocess.run([net_test_exe, "shards"], input="".join(f"{t} {w}\n" for t, w in cases), capture_output=True, text=True, timeout=60, env=env)
    assert run.returncode == 0, run.stdout + run.stderr
    lines = run.stdout.strip().split("\n")
    assert len(lines) == len(cases)
    for (t, w), line in zip(cases, lines):
        got = [int(x) for x in line.split()]
        assert list(zip(got[0::2], got[1::2])) == dp.all_shards(t, w), (t, w)


@pytest.mark.parametrize("source", ["net_cpp.cc", "rosenbrock_cpp.cc"])
def test_device_instantiation_compiles_and_fails_loudly_without_a_gpu(source):
    """examples/net_cpp.cc (the network layer over the device algebra, gadcu::Comm) and examples/rosenbrock_cpp.cc (one generic
    program on Eval, Graph1 and gadcu::BatchedGraph1) build warning-free against the mirror headers."""
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "example")
        res = build_device_example(exe, source)
        assert res.returncode == 0, res.stderr
        try:
            import torch
            has_gpu = torch.cuda.is_available()
        except Exception:
            has_gpu = False
        if not has_gpu:
            run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
            assert run.returncode == 2 and "gadcu::Error kind 8" in run.stdout, run.stdout  # Error{Cuda}: no CPU fallback
