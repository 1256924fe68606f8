#!/usr/bin/env python
"""Per-kernel SASS opcode census of the shipped library and of one specialised (JIT) gate-group kernel.

    python tools/sass_census.py > profiles/r02_sass_census.md

Evidence for which hardware paths the kernels use (B200_PROFILING.md: UBLKCP / UTMALDG = TMA, DMMA = FP64 tensor cores,
LDGSTS = cp.async, SYNCS = mbarrier, BAR = CTA barriers).  Runs without a GPU (cuobjdump + NVRTC)."""
import collections
import ctypes
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
KEY = ["UBLKCP", "UTMALDG", "UTMASTG", "UBLKPF", "LDGSTS", "SYNCS", "BAR", "DMMA", "DFMA", "DADD", "DMUL", "LDS", "STS", "LDG", "STG",
       "LDL", "STL", "SHFL", "BRA", "BRX", "IMAD", "LOP3", "UMOV", "ISETP", "FSEL", "R2UR"]


def census(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    kernels, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    return kernels


def table(kernels, keep=None):
    rows = []
    for name, c in sorted(kernels.items(), key=lambda kv: -sum(kv[1].values())):
        if keep and not any(k in name for k in keep):
            continue
        total = sum(c.values())
        short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:70]
        cells = " ".join(f"{k}:{c[k]}" for k in KEY if c[k])
        rows.append(f"| `{short}` | {total} | {cells} |")
    return "\n".join(["| kernel | SASS instructions | opcodes of interest |", "|---|---|---|"] + rows)


def main():
    lib = os.path.join(ROOT, "hyquas_b200", "libhyquas_b200.so")
    print("# SASS opcode census (round 2)\n")
    print(f"`cuobjdump -sass {os.path.relpath(lib, ROOT)}` (sm_100a), per kernel:\n")
    print(table(census(lib)))
    # one specialised kernel: the largest gate group of supremacy_24 as the product's own partitioner cuts it (host-only
    # compile, no GPU; HQ_JIT_DUMP_DIR keeps every emitted kernel source)
    from hyquas_b200 import api, circuits as C
    from hyquas_b200._lib import check, lib as L
    with tempfile.TemporaryDirectory() as dump:
        os.environ["HQ_JIT_DUMP_DIR"] = dump
        os.environ["HQ_BACKEND"] = "group"
        api.init_host_only(1, 0)
        c = api.Circuit.from_qasm(C.generate("supremacy_24"))
        c.compile()
        groups = c.groups()
        big = max(range(len(groups)), key=lambda i: groups[i]["gates"])
        rounds, fp = ctypes.c_int(), ctypes.c_double()
        check(L.hq_circuit_group_cost(c._h, big, rounds, fp))
        src = open(os.path.join(dump, "group_%03d.cu" % big)).read()
        ngates = groups[big]["gates"]
        c.close()
    class _Buf:
        value = src.encode()
    buf = _Buf()
    keep = [None] * ngates
    with tempfile.TemporaryDirectory() as d:
        cubin = os.path.join(d, "k.cubin")
        log = ctypes.create_string_buffer(1 << 16)
        check(L.hq_debug_jit_compile_to_file(buf.value, cubin.encode(), log, len(log)))
        print(f"\nOne specialised gate-group kernel (NVRTC, sm_100a): supremacy-like group, {len(keep)} gates, {rounds.value} rounds, "
              f"{fp.value:.1f} FP64 instructions per amplitude by the emitter's count:\n")
        print(table(census(cubin)))
        regs = re.search(r"Used (\d+) registers", log.value.decode())
        spill = re.search(r"(\d+) bytes spill stores", log.value.decode())
        print(f"\nptxas: {regs.group(1) if regs else '?'} registers, {spill.group(1) if spill else '0'} bytes of spill stores.")


if __name__ == "__main__":
    main()
