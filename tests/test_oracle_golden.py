"""The oracle against the reference's golden vectors and against an independent numpy replay (CPU only)."""
import hashlib
import json
import math
import os

import numpy as np
import pytest

from hyquas_b200 import circuits as C
from oracle import oracle as O


def _sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def test_goldens_are_the_reference_files(golden_dir):
    """tests/golden/* are byte-exact reconstructions: sha256 and size equal the Git-LFS pointer of the reference."""
    oids = json.load(open(os.path.join(golden_dir, "lfs_oids.json")))["files"]
    for ours, theirs in [("qft_28.qasm", "tests/input/qft_28.qasm"), ("qft_28.log", "tests/output/qft_28.log"),
                         ("bv_28.qasm", "tests/input/bv_28.qasm"), ("bv_28.log", "tests/output/bv_28.log"),
                         ("hidden_shift_28.log", "tests/output/hidden_shift_28.log")]:
        p = os.path.join(golden_dir, ours)
        assert _sha(p) == oids[theirs]["sha256"], ours
        assert os.path.getsize(p) == oids[theirs]["size"], ours


def test_generators_reproduce_reference_inputs(golden_dir):
    assert C.qft(28) == open(os.path.join(golden_dir, "qft_28.qasm")).read()
    assert C.bv(28) == open(os.path.join(golden_dir, "bv_28.qasm")).read()


@pytest.mark.parametrize("seed", range(4))
def test_oracle_c_matches_numpy_replay(seed):
    n, gates = O.parse_qasm(C.random_circuit(10, 300, seed, names=None))
    a, b = O.simulate(n, gates), O.simulate_numpy(n, gates)
    assert np.max(np.abs(a - b)) < 1e-14
    assert abs(np.vdot(a, a).real - 1.0) < 1e-12


def test_parser_rules():
    n, g = O.parse_qasm('OPENQASM 2.0;\ninclude "qelib1.inc";\n// comment\nqreg q[9];\ncu1(pi/4) q[0],q[3];\n'
                        'u3(pi*0.5,0.25,-1.5) q[2]; // trailing text is dropped\nccx q[1],q[2],q[8];\n')
    assert n == 9 and len(g) == 3
    assert (g[0].control, g[0].target) == (0, 3) and g[0].params == (math.acos(-1) / 4,)
    assert g[1].params == (math.acos(-1) * 0.5, 0.25, -1.5)
    assert (g[2].control, g[2].control2, g[2].target) == (1, 2, 8)
    with pytest.raises(SystemExit):
        O.parse_qasm("qreg q[3];\nswap q[0],q[1];\n")


def test_dump_format_and_zero_wrapper():
    s = np.zeros(256, dtype=np.complex128)
    s[3] = complex(-1e-15, 0.6)
    s[200] = 0.8
    text = O.dump_state(s, 8)
    lines = text.splitlines()
    assert len(lines) == 129
    assert lines[3] == "3 0.360000000000: 0.000000000000 0.600000000000"   # -1e-15 prints as +0
    assert lines[128] == "200 0.640000000000: 0.800000000000 0.000000000000"


@pytest.mark.parametrize("n", [12, 16, 20])
def test_families_small_analytic(n):
    """Same closed forms as the n=28 goldens: qft -> uniform 2^-n/2, bv -> (|s,0> - |s,1>)/sqrt2, hidden_shift -> |s>."""
    _, st = O.simulate_qasm(C.qft(n))
    assert np.max(np.abs(st - 2.0 ** (-n / 2))) < 1e-12
    _, st = O.simulate_qasm(C.bv(n))
    lo, hi = (1 << (n - 1)) - 1, (1 << n) - 1
    assert abs(st[lo] - 1 / math.sqrt(2)) < 1e-12 and abs(st[hi] + 1 / math.sqrt(2)) < 1e-12
    shift = 0x2B5A5 & ((1 << n) - 1)
    _, st = O.simulate_qasm(C.hidden_shift(n, shift))
    assert abs(st[shift] - 1.0) < 1e-12


def test_adder_adds():
    n = 14
    text = C.adder(n)
    _, gates = O.parse_qasm(text)
    st = O.simulate(n, gates)
    idx = int(np.argmax(np.abs(st)))
    assert abs(abs(st[idx]) - 1) < 1e-12
    m = (n - 2) // 2
    xs = [g.target for g in gates if g.name == "x"]
    a = sum(1 << i for i in range(m) if (2 + 2 * i) in xs)
    b = sum(1 << i for i in range(m) if (1 + 2 * i) in xs)
    total = sum(((idx >> (1 + 2 * i)) & 1) << i for i in range(m)) + (((idx >> (n - 1)) & 1) << m)
    assert total == a + b


@pytest.mark.slow
@pytest.mark.parametrize("name", ["qft_28", "bv_28", "hidden_shift_28"])
def test_oracle_reproduces_reference_goldens_28(name, golden_dir):
    """Byte-exact against tests/output/*.log of the reference (same check as oracle/pin_goldens.py)."""
    path = os.path.join(golden_dir, name + ".qasm")
    text = open(path).read() if os.path.exists(path) else C.hidden_shift(28)
    n, st = O.simulate_qasm(text)
    assert O.dump_state(st, n) == open(os.path.join(golden_dir, name + ".log")).read()
