"""GPU: the reference's tests replayed through the C++ host mirror (include/gadcu.hpp -> C ABI -> device backend):
gad/tests/graph.rs:8-69, gad/tests/extension.rs:61-107, Check / Error variants, matmul VJP on exact integers, WeightOps
under an apply_gradient_step-style loop. The program is examples/reference_tests.cc; it exits 0 when every expectation
holds."""
import os
import shutil
import subprocess
import tempfile

import pytest

from test_check_and_abi import _build_cpp_reference_tests

pytestmark = pytest.mark.gpu


def test_reference_tests_through_the_cpp_mirror():
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "reference_tests")
        res = _build_cpp_reference_tests(exe)
        assert res.returncode == 0, res.stderr
        run = subprocess.run([exe], capture_output=True, text=True, timeout=300)
        assert run.returncode == 0, run.stdout + run.stderr
        assert "all reference tests passed" in run.stdout
