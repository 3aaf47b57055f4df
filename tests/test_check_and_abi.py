"""CPU: the product's Check algebra (through the C ABI, no GPU needed) against the oracle's, bit-exact
incl. the error variant; and the shape of the C ABI itself."""
import itertools
import os
import re
import subprocess

import pytest

import gad_b200 as gb
from gad_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DIMS = [(1, 1, 1, 1), (4, 3, 1, 1), (3, 4, 1, 1), (3, 5, 1, 1), (4, 1, 1, 1), (1, 3, 1, 1), (12, 1, 1, 1), (2, 2, 3, 1),
        (4, 3, 2, 2), (3, 5, 2, 2), (3, 5, 7, 1), (4, 3, 2, 1), (3, 5, 3, 1), (2, 3, 1, 1), (8, 6, 1, 1)]


def _both(oracle, name, args_dims, extra=()):
    """Run one Check op on both sides; return ((kind or None, dims), (kind or None, dims))."""
    chk = gb.Check()
    oc = oracle.OCheck()
    try:
        got = (None, getattr(chk, name)(*args_dims, *extra))
    except gb.GadError as e:
        got = (e.kind, None)
    try:
        ovals = [oc.dims_value(d) if d is not None else None for d in args_dims]
        want = (None, getattr(oc, name)(*ovals, *extra).dims())
    except oracle.OracleError as e:
        want = (e.kind, None)
    return got, want


@pytest.mark.parametrize("name", ["add", "sub", "mul", "div", "pow", "min", "max", "dot"])
def test_check_binary_rules(oracle, name):
    for a, b in itertools.product(DIMS, DIMS):
        got, want = _both(oracle, name, [a, b])
        assert got == want, (name, a, b, got, want)


@pytest.mark.parametrize("name", ["moddims", "tile_as", "sum_as", "max_as", "argmax_as", "softmax_as"])
def test_check_shape_rules(oracle, name):
    for v, r in itertools.product(DIMS, DIMS):
        got, want = _both(oracle, name, [v], extra=(r,))
        assert got == want, (name, v, r, got, want)


def test_check_unary_and_matrix_rules(oracle):
    for v in DIMS:
        for name in ["flat", "as_scalar", "transpose", "exp", "neg", "zeros", "relu", "abs"]:
            got, want = _both(oracle, name, [v])
            assert got == want, (name, v, got, want)
    for a, b in itertools.product(DIMS, DIMS):
        for t1, t2 in itertools.product([False, True], [False, True]):
            got, want = _both(oracle, "matmul", [a, b], extra=(gb.MatProp(t1), gb.MatProp(t2)))
            assert got == want, ("matmul", a, b, t1, t2, got, want)
        got, want = _both(oracle, "select_argmax", [a, b, a, b])
        assert got == want
        got, want = _both(oracle, "select_argmax", [a, a, b, None])
        assert got == want


def test_check_scalars_always_ok(oracle):
    chk = gb.Check()
    assert chk.add((), ()) == () and chk.mul((), ()) == () and chk.exp(()) == () and chk.select_argmax((), (), (), ()) == ()
    with pytest.raises(gb.EmptyError):
        chk.add_all([])


def test_error_variants_match_reference_enum():
    """gad/src/error.rs:9-40 order: Dimensions, ReducedDimensions, Lengths, Empty, MissingId,
    MissingGradient, MissingNode."""
    chk = gb.Check()
    with pytest.raises(gb.DimensionsError) as e:
        chk.sum_as((4, 3, 1, 1), (2, 3))
    assert e.value.status == 1
    with pytest.raises(gb.ReducedDimensionsError) as e:
        chk.max_as((4, 3, 1, 1), (2, 3))
    assert e.value.status == 2
    with pytest.raises(gb.MissingIdError):
        gb.Value(None, None).gid()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gadcu.h")).read()
    return sorted(set(re.findall(r"GADCU_API\s+[\w\s\*]+?\b(gadcu_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) > 90
    exported = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\b(gadcu_\w+)\b", exported))
    missing = [n for n in names if n not in exported]
    assert not missing, missing
    unbound = [n for n in names if n not in _lib.SIGNATURES]
    assert not unbound, unbound
    for n in names:
        assert getattr(lib, n) is not None
    # nothing but the C ABI leaks out of the library
    leaked = [s for s in exported if not s.startswith("gadcu_")]
    assert not leaked


def _split_params(params: str):
    """Top-level comma split of a C or Rust parameter list (function-pointer parameters contain commas of their own)."""
    out, depth, cur = [], 0, ""
    for ch in params:
        if ch in "(<[":
            depth += 1
        elif ch in ")>]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip() and cur.strip() != "void":
        out.append(cur.strip())
    return out


def test_rust_binding_declares_the_whole_header_with_matching_arity():
    """rust/src/ffi.rs is what a gad maintainer would compile against libgadcu; there is no Rust
    toolchain here, so at least keep it in step with include/gadcu.h: same entry points, same number of parameters,
    same status codes."""
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "gadcu.h")).read(), flags=re.S)
    rust = re.sub(r"//[^\n]*", "", open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read())
    c_fns = {}
    for m in re.finditer(r"GADCU_API\s+[\w\s\*]+?\b(gadcu_\w+)\s*\(", header):
        start, depth, i = m.end(), 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(header[i], 0)
            i += 1
        c_fns[m.group(1)] = _split_params(header[start:i - 1])
    r_fns = {}
    for m in re.finditer(r"pub fn (gadcu_\w+)\s*\(", rust):
        start, depth, i = m.end(), 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(rust[i], 0)
            i += 1
        r_fns[m.group(1)] = _split_params(rust[start:i - 1])
    assert sorted(c_fns) == sorted(r_fns), sorted(set(c_fns) ^ set(r_fns))
    wrong = {n: (len(c_fns[n]), len(r_fns[n])) for n in c_fns if len(c_fns[n]) != len(r_fns[n])}
    assert not wrong, wrong
    c_codes = dict(re.findall(r"(GADCU_ERR_\w+)\s*=\s*(\d+)", header))
    r_codes = dict(re.findall(r"pub const (GADCU_ERR_\w+): gadcu_status = (\d+);", rust))
    assert c_codes == r_codes


def test_no_compute_without_gpu_is_an_error_not_a_fallback():
    """On a host without a GPU, creating a context fails with a CUDA error (status 8): the product
    never silently computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(gb.CudaError):
        gb.Context(0)


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under gad_b200/ may import, link or name it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gad_b200")):
        if "build" in dirpath.split(os.sep)[-1:] or dirpath.endswith("lib"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cc")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"oracle_py|liboracle|gad_oracle|oracle/", text):
                    if f == "__init__.py" and "nothing here imports" in text:
                        continue
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
    deps = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in deps


def test_header_is_plain_c_and_the_c_demo_links():
    """include/gadcu.h is usable from C (no C++-isms) and examples/c_abi_demo.c links against libgadcu.so."""
    import shutil
    import tempfile
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as tmp:
        res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", os.path.join(root, "examples", "c_abi_demo.c"), "-I", os.path.join(root, "include"),
                              "-L", os.path.dirname(_lib.LIB_PATH), "-lgadcu", "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH),
                              "-o", os.path.join(tmp, "demo")], capture_output=True, text=True)
        assert res.returncode == 0, res.stderr


def _build_cpp_reference_tests(out_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    return subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", os.path.join(root, "examples", "reference_tests.cc"),
                           "-I", os.path.join(root, "include"), "-L", os.path.dirname(_lib.LIB_PATH), "-lgadcu",
                           "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH), "-o", out_path], capture_output=True, text=True)


def test_cpp_host_mirror_compiles_and_fails_loudly_without_a_gpu():
    """include/gadcu.hpp (the compiled-language mirror of gad's traits above the C ABI) builds warning-free, and the
    reference tests written against it (examples/reference_tests.cc) link against libgadcu.so. Without a GPU the
    program reports Error{Cuda} and exits non-zero: no CPU fallback."""
    import shutil
    import tempfile
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "reference_tests")
        res = _build_cpp_reference_tests(exe)
        assert res.returncode == 0, res.stderr
        try:
            import torch
            has_gpu = torch.cuda.is_available()
        except Exception:
            has_gpu = False
        if not has_gpu:
            run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
            assert run.returncode == 2 and "gadcu::Error kind 8" in run.stdout, run.stdout


def test_tape_node_container_semantics():
    """gadcu::SmallVec (csrc/algebra.h), the inline-first container of a node's inputs and captured values: copy, move,
    growth past the inline capacity, assignment in every direction and `Node::clear`, checked with an element type that
    counts its constructions and destructions (tests/cpp/smallvec_test.cc; host-only, no CUDA call)."""
    import shutil
    import tempfile
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("no CUDA headers")
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "smallvec_test")
        res = subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-fsanitize=address,undefined", "-I", os.path.join(root, "gad_b200", "csrc"),
                              "-I", os.path.join(root, "include"), "-I", cuda_inc, os.path.join(root, "tests", "cpp", "smallvec_test.cc"), "-o", exe],
                             capture_output=True, text=True)
        if res.returncode != 0 and "sanitize" in res.stderr:  # (no sanitizer runtime in the image: the counting element type still checks)
            res = subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-I", os.path.join(root, "gad_b200", "csrc"), "-I", os.path.join(root, "include"),
                                  "-I", cuda_inc, os.path.join(root, "tests", "cpp", "smallvec_test.cc"), "-o", exe], capture_output=True, text=True)
        assert res.returncode == 0, res.stderr
        run = subprocess.run([exe], capture_output=True, text=True)
        assert run.returncode == 0 and run.stdout.strip().endswith("ok"), run.stdout + run.stderr
