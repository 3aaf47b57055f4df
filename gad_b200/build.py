"""Build the gadcu shared library (hand-written CUDA for sm_100a) in-tree.

    python -m gad_b200.build            # incremental
    python -m gad_b200.build --force

nvcc cross-compiles without a GPU. The result, gad_b200/lib/libgadcu.so, is git-ignored but travels
with the repo snapshot to the GPU box. Elementwise / reduction / lane-interpreter translation units
are compiled with -fmad=false so that every tape operation rounds on its own (bit-parity between the
fused and the unfused reverse sweep, and op-for-op agreement with the CPU oracle's op order).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libgadcu.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "-I", os.path.join(ROOT, "include"), "-I", CSRC, "--expt-relaxed-constexpr", "-Xcudafe", "--diag_suppress=177"]
NO_FMA = {"ew.cu", "reduce.cu", "batched.cu", "gemm_tc.cu"}  # gemm_tc.cu: its fused elementwise epilogues must round like the tape's separate kernels
SOURCES = ["runtime.cu", "eval.cu", "graph.cu", "capi.cu", "ew.cu", "reduce.cu", "gemm_simt.cu", "gemm_tc.cu", "gemm_dmma.cu",
           "batched.cu", "comm.cu", "jit.cu", "symm.cu"]


def _headers_mtime() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    paths.append(os.path.join(ROOT, "include", "gadcu.h"))
    return max(os.path.getmtime(p) for p in paths)


def _compile(src: str, force: bool, hdr_mtime: float) -> str:
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(path)
            and os.path.getmtime(obj) > hdr_mtime):
        return obj
    cmd = [NVCC, *ARCH, *COMMON, "-c", path, "-o", obj]
    if src in NO_FMA:
        cmd.insert(1, "-fmad=false")
    if os.environ.get("GADCU_PTXAS_V"):
        cmd += ["-Xptxas", "-v"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
    if os.environ.get("GADCU_PTXAS_V"):
        sys.stderr.write(res.stderr)
    return obj


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    hdr = _headers_mtime()
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, hdr), SOURCES))
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-ldl", "-lpthread"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(f"built {LIB}")
    return LIB


def build_examples(verbose: bool = True) -> list:
    """Compile the stand-alone C hosts under examples/ against the built library (gcc; they only run on a GPU box).
    examples/c2_host is what bench.py times for the API-level batched-tape figure: a compiled host recording its tape
    through the C ABI every step, as a Rust gad program would."""
    out = []
    lib_dir = os.path.relpath(LIBDIR, os.path.join(ROOT, "examples"))
    for name in ("c2_host",):
        src, exe = os.path.join(ROOT, "examples", name + ".c"), os.path.join(ROOT, "examples", name)
        if os.path.exists(exe) and os.path.getmtime(exe) > max(os.path.getmtime(src), os.path.getmtime(LIB)):
            out.append(exe)
            continue
        cmd = [os.environ.get("CC", "gcc"), "-O2", "-Wall", src, "-I", os.path.join(ROOT, "include"), "-L", LIBDIR, "-lgadcu",
               "-Wl,-rpath,$ORIGIN/" + lib_dir, "-o", exe]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"building {name} failed:\n{res.stdout}\n{res.stderr}")
        if verbose:
            print(f"built {exe}")
        out.append(exe)
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    build_examples()
