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
    # one specialised kernel: a supremacy-like group of ~100 gates on a 12-bit tile
    import numpy as np  # noqa: F401
    from hyquas_b200 import circuits as C
    from hyquas_b200._lib import HqGate, check, lib as L
    from oracle import oracle as O
    _, gates = O.parse_qasm(C.supremacy(20, cycles=10, seed=5))
    keep = [g for g in gates if (g.mat[0, 1] == 0 and g.mat[1, 0] == 0) or g.target < 12]
    arr = (HqGate * len(keep))()
    for i, g in enumerate(keep):
        arr[i].type, arr[i].target, arr[i].control, arr[i].control2 = 0, g.target, g.control, g.control2
        m = g.mat.reshape(4)
        for j in range(4):
            arr[i].mat[2 * j], arr[i].mat[2 * j + 1] = m[j].real, m[j].imag
    plan = ctypes.c_void_p()
    check(L.hq_group_plan_create(20, 0xFFF, arr, len(keep), ctypes.byref(plan)))
    need = ctypes.c_size_t()
    check(L.hq_debug_group_plan_jit_source(plan, 0, None, 0, ctypes.byref(need)))
    buf = ctypes.create_string_buffer(need.value)
    check(L.hq_debug_group_plan_jit_source(plan, 0, buf, need.value, ctypes.byref(need)))
    rounds, fp = ctypes.c_int(), ctypes.c_double()
    check(L.hq_group_plan_cost(plan, rounds, fp))
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
