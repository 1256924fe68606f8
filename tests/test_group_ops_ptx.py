"""The gate-group kernel's arithmetic is generated inline PTX (tools/gen_group_ops.py).  This test interprets every
generated body (a handful of FP64 opcodes on named registers) on random data and checks it against the 2x2 matrix the
planner means by that op, so sign/pairing mistakes in the generator are caught on a machine without a GPU.  It also
checks that the checked-in .inc files are what the generator produces."""
import importlib.util
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("gen_group_ops", os.path.join(ROOT, "tools", "gen_group_ops.py"))
G = importlib.util.module_from_spec(spec)
spec.loader.exec_module(G)

NEG_BITS = np.uint64(0x8000000000000000)


def run_ptx(lines, regs, creg, m):
    """regs: dict name -> float.  Supports the opcode subset the generator emits."""
    def val(tok):
        tok = tok.strip()
        if tok.startswith("%"):
            assert tok == "%1"
            return float(creg)
        if tok.startswith("0d"):
            return float(np.array([int(tok[2:], 16)], dtype=np.uint64).view(np.float64)[0])
        if tok.startswith("0x"):
            return int(tok, 16)
        if re.fullmatch(r"-?\d+", tok):
            return int(tok)
        return regs[tok]

    pred = {}
    for raw in lines:
        line = raw.strip().rstrip(";")
        guard = True
        if line.startswith("@"):
            p, line = line.split(" ", 1)
            guard = pred[p[1:]]
        op, rest = line.split(" ", 1)
        if op == "ld.shared.v2.f64":   # {hqm2j, hqm2j+1}, [%2+16j]: the op's coefficients, as the planner stored them
            mm = re.fullmatch(r"\{(hqm\d), (hqm\d)\}, \[%2\+(\d+)\]", rest)
            j = int(mm.group(3)) // 8
            assert mm.group(1) == f"hqm{j}" and mm.group(2) == f"hqm{j + 1}" and guard
            regs[mm.group(1)], regs[mm.group(2)] = m[j], m[j + 1]
            continue
        args = [a.strip() for a in rest.split(",")]
        if not guard:
            continue
        d = args[0]
        if op == "mul.f64":
            regs[d] = val(args[1]) * val(args[2])
        elif op == "fma.rn.f64":   # exact enough for a 1e-13 check
            regs[d] = val(args[1]) * val(args[2]) + val(args[3])
        elif op == "add.f64":
            regs[d] = val(args[1]) + val(args[2])
        elif op == "sub.f64":
            regs[d] = val(args[1]) - val(args[2])
        elif op == "neg.f64":
            regs[d] = -val(args[1])
        elif op == "mov.f64":
            regs[d] = val(args[1])
        elif op == "xor.b64":
            assert val(args[2]) == int(NEG_BITS)
            regs[d] = -val(args[1])
        elif op == "and.b32":
            a = int(creg) if args[1] == "%1" else int(val(args[1]))
            regs[d] = a & int(val(args[2]))
        elif op == "setp.eq.u32":
            pred[d] = int(val(args[1])) == int(val(args[2]))
        else:
            raise AssertionError(f"unexpected PTX opcode in generated body: {raw}")


def matrix_of(kind, m):
    """The 2x2 the planner means by (kind, coefficients m[0..7])."""
    c = lambda i: complex(m[i], m[i + 1])
    if kind == "GEN":       # the complex matrix itself
        return np.array([[c(0), c(2)], [c(4), c(6)]])
    if kind == "REAL":      # {a, b, c, d} real
        return np.array([[m[0], m[1]], [m[2], m[3]]], dtype=complex)
    if kind == "RXL":       # {a, b, c, d} of [[a, i b], [i c, d]]
        return np.array([[m[0], 1j * m[1]], [1j * m[2], m[3]]])
    if kind == "SWAP":
        return np.array([[0, 1], [1, 0]], dtype=complex)
    if kind == "YL":
        return np.array([[0, -1j], [1j, 0]])
    if kind == "DIAG_R":
        return np.diag([c(0), c(6)])
    if kind == "DIAG_R1":
        return np.diag([1, c(6)])
    if kind == "ZFLIP":
        return np.diag([1, -1]).astype(complex)
    if kind.startswith("BF"):
        p, q = G.BF_PQ[int(kind[2:])]
        return np.array([[1, p], [q, -p * q]], dtype=complex)
    raise AssertionError(kind)


@pytest.mark.parametrize("rbits", [3, 4])
def test_every_generated_body_matches_its_matrix(rbits):
    R = 1 << rbits
    rng = np.random.default_rng(rbits)
    cat = G.catalog(rbits)
    assert len({code for code, *_ in cat}) == len(cat)
    for code, kind, tb, cbc, lines in cat:
        assert code == G.KINDS.index(kind) * 24 + tb * 6 + cbc
        cregs = [0] if cbc == 0 else ([1 << (cbc - 1)] if cbc <= 4 else
                                       [x for x in range(R) if not (x >> tb & 1)])   # generic: any mask avoiding the target
        for creg in cregs:
            m = rng.standard_normal(8)
            amps = rng.standard_normal(R) + 1j * rng.standard_normal(R)
            regs = {}
            for i in range(R):
                regs[f"hqa{2 * i}"], regs[f"hqa{2 * i + 1}"] = amps[i].real, amps[i].imag
            run_ptx(lines, regs, creg, m)
            got = np.array([complex(regs[f"hqa{2 * i}"], regs[f"hqa{2 * i + 1}"]) for i in range(R)])
            U = matrix_of(kind, m)
            want = amps.copy()
            for lo in range(R):
                if lo >> tb & 1 or (lo & creg) != creg:
                    continue
                hi = lo | 1 << tb
                want[lo], want[hi] = U[0, 0] * amps[lo] + U[0, 1] * amps[hi], U[1, 0] * amps[lo] + U[1, 1] * amps[hi]
            assert np.max(np.abs(got - want)) < 1e-12, (kind, tb, cbc, creg)


@pytest.mark.parametrize("rbits", [3, 4])
def test_checked_in_inc_is_current(rbits):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_group_ops.py"), str(rbits)],
                         capture_output=True, text=True, check=True).stdout
    with open(os.path.join(ROOT, "hyquas_b200", "csrc", "device", f"group_ops_gen_r{rbits}.inc")) as f:
        assert f.read() == out
