"""CPU: what the built library is made of, read from its SASS with cuobjdump — no GPU involved (nvcc cross-compiles).

The claims DESIGN.md §3 makes about each hot kernel are checked against the machine code that ships to the GPU box:
the f32 GEMM is tcgen05 + TMA + TMEM (UTCHMMA / UTMALDG / LDTM, cta_group::2, multicast commit), the f64 GEMM is DMMA, the
data-parallel exchange reduces inside the NVSwitch (LDGMC = multimem.ld_reduce), the elementwise / reduction kernels move
128-bit vectors, the translation units whose results must round op by op carry no fused multiply-add, and no kernel spills.
The mnemonics are the ones /opt/skills/guides/B200_PROFILING.md names as proof of the hardware path."""
import collections
import functools
import re
import shutil
import subprocess

import pytest

from gad_b200 import _lib

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
pytestmark = pytest.mark.skipif(shutil.which(CUOBJDUMP) is None, reason="cuobjdump not found")


@functools.lru_cache(maxsize=None)
def kernels():
    """{demangled name: (Counter of full SASS opcodes, registers, local bytes)} for every kernel of libgadcu.so."""
    _lib.load()  # (raises if the library was not built)
    sass = subprocess.run([CUOBJDUMP, "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    res = subprocess.run([CUOBJDUMP, "--dump-resource-usage", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    usage, cur = {}, None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?LOCAL:(\d+)", line)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)))
    ops, cur = collections.defaultdict(collections.Counter), None
    opcode = re.compile(r"^\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z][A-Za-z0-9_.]*)")
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = opcode.match(line)
        if m and cur:
            ops[cur][m.group(1)] += 1
    names = list(ops)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
    return {d: (ops[n], *usage.get(n, (None, None))) for n, d in zip(names, dem)}


def named(part):
    found = {n: v for n, v in kernels().items() if part in n}
    assert found, f"no kernel named *{part}* in the library"
    return found


def count(ops, prefix):
    return sum(v for k, v in ops.items() if k.startswith(prefix))


def test_no_kernel_spills():
    spilling = {n: loc for n, (_, _, loc) in kernels().items() if loc}
    # the lane interpreter keeps its 192 value slots per lane in local memory by design (the fallback without NVRTC)
    spilling = {n: l for n, l in spilling.items() if "lane_interpreter" not in n}
    assert not spilling, spilling
    assert len(kernels()) > 100  # (the whole library was read, not a stub)


def test_f32_gemm_is_tcgen05_tma_tmem():
    for name, (ops, regs, _) in named("gemm_tf32x3_pair_kernel").items():
        assert count(ops, "UTCHMMA.2CTA") >= 6, (name, "tcgen05.mma cta_group::2")          # 3 MMAs per k-step, unrolled
        assert count(ops, "UTCHMMA") == count(ops, "UTCHMMA.2CTA"), name                   # every MMA is a pair MMA
        assert count(ops, "UTMALDG.2D") + count(ops, "UTMALDG.3D") >= 8, (name, "TMA tile loads")
        assert count(ops, "LDTM") >= 1, (name, "tcgen05.ld of the accumulator")
        assert any(k.startswith("UTCBAR") and "MULTICAST" in k for k in ops), (name, "tcgen05.commit multicast to both CTAs")
        assert count(ops, "UTCATOMSWS") >= 2, (name, "TMEM alloc / dealloc")
        assert count(ops, "HMMA") == 0, (name, "no warp-level MMA beside the tcgen05 path")  # (its FFMAs are tanhf's own polynomial)
        assert regs <= 168, (name, regs)  # 384 threads per CTA: more than 170 registers would not launch
    for name, (ops, _, _) in named("gemm_tf32x3_kernel").items():  # the one-CTA v1 kernel kept for A/B
        assert count(ops, "UTCHMMA") >= 6 and count(ops, "UTMALDG") >= 8 and count(ops, "LDTM") >= 1, name


def test_fused_epilogue_lives_in_the_tensor_core_kernel():
    """tanh (MUFU) and the TF32 copies are produced by the kernel that issues the MMAs — not by a pass after it."""
    for name, (ops, _, _) in named("gemm_tf32x3_pair_kernel").items():
        assert count(ops, "MUFU") >= 1 and count(ops, "UTCHMMA") >= 6, name
        # per accumulator column a warp stores 32 consecutive rows = 128 contiguous bytes: result, hi and lo => >= 3 x 32 stores
        assert count(ops, "STG.E") >= 96, (name, "result, hi and lo stored from the epilogue")


def test_f64_gemm_is_dmma():
    for name, (ops, _, _) in named("gemm_dmma_kernel").items():
        assert count(ops, "DMMA") >= 128, name
        assert count(ops, "DFMA") == 0, (name, "no CUDA-core product")


def test_exchange_reduces_in_the_switch():
    for name, (ops, _, _) in named("dp_nvls_allreduce_kernel").items():
        assert any(k.startswith("LDGMC") and "ADD" in k and "F32" in k for k in ops), (name, "multimem.ld_reduce.add.f32")
        assert count(ops, "STG.E.128") >= 1, (name, "multimem.st of the sum")
        assert count(ops, "ATOM") == 0 and count(ops, "RED") == 0, (name, "no atomics: the order of the sum is the fabric's, not a race")


def test_rounding_units_carry_no_fma():
    """ew.cu / reduce.cu / batched.cu / gemm_tc.cu are compiled with -fmad=false: every tape op rounds on its own, which is
    what makes the fused sweep equal the unfused one and the oracle's op order. An FFMA / DFMA in one of their
    kernels that is not part of a libdevice routine would break that."""
    libdevice = ("FExp", "FLog", "FSin", "FCos", "FTanh", "FSigmoid", "FPow", "FSqrt", "FReciprocal", "FDiv", "FNorm", "FSoftmax",
                 "VExp", "VLog", "VSin", "VCos", "VTanh", "VSigmoid", "VPow", "VSqrt", "VReciprocal", "VDiv", "FLog1p", "VLog1p", "FScale")
    for functor in ("FAdd", "FSub", "FMul", "FNeg", "FMax", "FMin", "FRelu", "FAbs", "FSign", "VMul", "VSub", "FAddC", "FMulC", "FSetC"):
        for name, (ops, _, _) in kernels().items():
            if f"::{functor}," in name and "ew_kernel" in name:
                assert count(ops, "FFMA") == 0 and count(ops, "DFMA") == 0, name
    assert any(f"::{f}," in n for n in kernels() for f in libdevice)  # (the naming the test relies on still holds)
    for name, (ops, _, _) in named("reduce_contig_kernel").items():
        assert count(ops, "FFMA") == 0 and count(ops, "DFMA") == 0, name


def test_streaming_kernels_use_128_bit_accesses():
    def wide(ops):
        return sum(v for k, v in ops.items() if k.startswith("LDG") and ".128" in k)
    for part in ("ew_kernel<float", "split_tf32_kernel"):
        for name, (ops, _, _) in named(part).items():
            assert wide(ops) >= 1, (name, dict(ops))
    # reductions come in a vector walk (aligned segments) and a scalar one (ragged segments): the vector walk is 128-bit
    for dt in ("float", "double"):
        assert any(wide(ops) >= 4 for ops, _, _ in named(f"reduce_contig_kernel<{dt}").values()), dt
