"""Static evidence from the built library (no GPU needed): per kernel, the SASS mnemonics that show which hardware path it
uses (B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG/UBLKCP, mma.sync f64 -> DMMA) and its
registers / spills (cuobjdump --dump-resource-usage).   python profiles/sass_evidence.py > profiles/r01_sass_evidence.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gad_b200", "lib", "libgadcu.so")
MNEMONICS = r"\b(LDGMC|UTC[A-Z]*MMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UTMALDG|UTMASTG|UBLKCP|HMMA|DMMA|SYNCS|ATOMG|REDG|RED|MEMBAR|LDG|STG|LDS|STS|SHFL|MUFU)\b"


def demangle(name):
    return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    usage, cur = {}, None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
    counts, cur = collections.defaultdict(collections.Counter), None
    pat = re.compile(MNEMONICS)
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur:
            for mn in pat.findall(line):
                counts[cur][mn] += 1
    print("cuobjdump -sass / --dump-resource-usage gad_b200/lib/libgadcu.so (sm_100a), one entry per kernel; static instruction counts")
    print("tensor-core path: UTCHMMA = tcgen05.mma (kind::tf32), UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM alloc/dealloc, LDTM = tcgen05.ld,")
    print("UTMALDG = cp.async.bulk.tensor (TMA), SYNCS = mbarrier ops, DMMA = mma.sync f64. In gemm_tf32x3_pair_kernel the STG group includes the")
    print("tile-pushing epilogue's stores to peer (NVLink-mapped) staging buffers, in the same kernel as the UTCHMMA tiles.")
    print("LDGMC.E.ADD.F32x4 = multimem.ld_reduce.add.v4.f32 (the NVSwitch reduces the address over all ranks); multimem.st is an STG.E.128.STRONG.SYS on the multicast address.\n")
    for fn in sorted(counts, key=lambda f: (-sum(v for k, v in counts[f].items() if k.startswith("UTC") or k in ("DMMA", "LDTM", "UTMALDG")), f)):
        name = demangle(fn)
        if "gadcu" not in name and not name.startswith(("dp_", "void dp_")) and "dp_nvls" not in name:
            continue  # (the exchange kernels of comm.cu sit in an anonymous namespace inside extern "C": no gadcu:: in their names)
        reg, sh, loc = usage.get(fn, (None, None, None))
        c = counts[fn]
        keys = [k for k in ("LDGMC", "UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "DMMA", "HMMA", "LDG", "STG", "ATOMG", "RED", "REDG", "MEMBAR", "SHFL", "MUFU") if c.get(k)]
        print(f"{name[:150]}\n    registers {reg}, static shared {sh} B, local (spills) {loc} B;  " + ", ".join(f"{k} x{c[k]}" for k in keys))


if __name__ == "__main__":
    main()
