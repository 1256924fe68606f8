"""The circuit-level peephole pass (hyquas_b200/csrc/host/peephole.cpp): cx-diag-cx and h-cx..cx-h patterns become diagonal
gates before partitioning.  The oracle always replays the ORIGINAL gate list, so these tests prove the rewrite keeps the
amplitudes (<= 1e-12 on the CPU plan emulator) -- including the near-miss orders where a rewrite would be wrong."""
import random

import numpy as np
import pytest

from hyquas_b200 import api, circuits as C
from hyquas_b200._lib import check, lib
from oracle import oracle as O


@pytest.fixture(autouse=True)
def _count_gates_without_the_merge_pass(monkeypatch, request):
    """The pattern tests below count gates in and out; the single-qubit merge pass (P4) has its own tests at the end."""
    if "merge" not in request.node.name:
        monkeypatch.setenv("HQ_PEEPHOLE_MERGE", "0")


def run(text):
    api.init_host_only(1, 0)
    c = api.Circuit.from_qasm(text)
    c.compile()
    n = c.num_qubits
    s = O.zero_state(n)
    check(lib.hq_debug_circuit_emulate(c._h, s.ctypes.data))
    info = c.schedule_info()
    c.close()
    _, gates = O.parse_qasm(text)
    return s, O.simulate(n, gates), info, len(gates)


def qasm(n, lines):
    return "OPENQASM 2.0;\ninclude \"qelib1.inc\";\nqreg q[%d];\n" % n + "".join(l + "\n" for l in lines)


PREP = ["h q[0];", "ry(0.3) q[1];", "rx(1.1) q[2];", "h q[3];", "u3(0.4,0.2,0.9) q[4];", "ry(2.0) q[5];", "h q[6];", "rx(0.7) q[7];",
        "ry(0.9) q[8];", "h q[9];"]


def test_zz_pattern_becomes_diagonal_and_matches():
    body = ["cx q[1],q[2];", "rz(0.37) q[2];", "cx q[1],q[2];",          # plain
            "cx q[3],q[4];", "h q[7];", "t q[4];", "ry(0.2) q[8];", "cx q[3],q[4];",   # other qubits interleaved, D = t
            "cx q[5],q[0];", "u1(1.3) q[0];", "cx q[5],q[0];"]
    got, want, info, ngates = run(qasm(10, PREP + body + ["h q[1];", "rx(0.3) q[4];", "h q[0];"]))
    assert np.max(np.abs(got - want)) < 1e-12
    assert info["gates"] == ngates          # 3 gates in, 3 diagonal gates out per pattern
    # the rewritten circuit needs no cx at all: one tile-kernel sweep is enough for 10 qubits
    assert info["groups"] == 1


@pytest.mark.parametrize("middle", [
    ["cx q[1],q[2];", "rz(0.5) q[2];", "cx q[1],q[2];"],                  # a ZZ term on (1,2) INSIDE a cx(0,1) ... cx(0,1) pair:
])
def test_nested_patterns_are_not_merged_wrongly(middle):
    """cx 0,1 ; [ZZ(1,2)] ; t 1 ; cx 0,1 -- after the inner rewrite nothing but diagonal gates sits between the outer cx pair,
    but more than ONE of them touches qubit 1, so the outer pair must stay (or be rewritten correctly): amplitudes decide."""
    body = ["cx q[0],q[1];"] + middle + ["t q[1];", "cx q[0],q[1];"]
    got, want, _, _ = run(qasm(10, PREP + body + ["h q[1];", "h q[2];"]))
    assert np.max(np.abs(got - want)) < 1e-12


def test_h_cx_h_becomes_cz():
    body = ["h q[9];", "cx q[0],q[9];", "rx(0.4) q[0];", "cx q[3],q[9];", "cx q[0],q[9];", "h q[9];"]
    got, want, info, ngates = run(qasm(10, PREP + body + ["h q[3];"]))
    assert np.max(np.abs(got - want)) < 1e-12
    assert info["gates"] == ngates - 2      # the two h are gone, the three cx are cz


def test_h_cx_h_blocked_by_other_gate_on_target():
    body = ["h q[9];", "cx q[0],q[9];", "z q[9];", "cx q[3],q[9];", "h q[9];"]
    got, want, info, ngates = run(qasm(10, PREP + body))
    assert np.max(np.abs(got - want)) < 1e-12
    assert info["gates"] == ngates


def test_x_diag_x_becomes_diagonal(monkeypatch):
    """x a ; diagonal gates touching a (as target, as control, uncontrolled) ; x a  ->  no x left (opt-in pass)."""
    monkeypatch.setenv("HQ_PEEPHOLE_X", "1")
    body = ["x q[3];", "cz q[3],q[5];", "h q[8];", "cu1(0.7) q[6],q[3];", "rz(0.9) q[3];", "t q[3];", "crz(1.1) q[3],q[7];", "x q[3];",
            "x q[9];", "cz q[0],q[9];", "x q[9];"]
    api.logger_flush()
    got, want, info, ngates = run(qasm(10, PREP + body + ["h q[3];", "h q[5];", "rx(0.2) q[9];"]))
    assert np.max(np.abs(got - want)) < 1e-12
    import re
    m = re.search(r"(\d+) x-diag-x patterns", api.logger_flush())
    assert m and int(m.group(1)) == 2


def test_x_diag_x_blocked_by_non_diagonal(monkeypatch):
    monkeypatch.setenv("HQ_PEEPHOLE_X", "1")
    body = ["x q[3];", "cz q[3],q[5];", "cx q[3],q[4];", "x q[3];",        # cx with control 3 is not diagonal: must stay
            "x q[6];", "h q[6];", "x q[6];"]
    got, want, info, ngates = run(qasm(10, PREP + body))
    assert np.max(np.abs(got - want)) < 1e-12
    assert info["gates"] == ngates


def test_hidden_shift_loses_its_x_gates(monkeypatch):
    monkeypatch.setenv("HQ_PEEPHOLE_X", "1")
    text = C.generate("hidden_shift_14")
    got, want, info, ngates = run(text)
    assert np.max(np.abs(got - want)) < 1e-12
    nx = sum(1 for l in text.splitlines() if l.startswith("x "))
    assert nx > 0 and info["gates"] <= ngates       # every x pair is gone; each cz in between became z + cz


def test_switch_off(monkeypatch):
    text = C.generate("qaoa_12")
    monkeypatch.setenv("HQ_PEEPHOLE", "0")
    got0, want, info0, _ = run(text)
    monkeypatch.delenv("HQ_PEEPHOLE")
    got1, _, info1, _ = run(text)
    assert np.max(np.abs(got0 - want)) < 1e-12 and np.max(np.abs(got1 - want)) < 1e-12
    assert info1["groups"] <= info0["groups"]


@pytest.mark.parametrize("seed", range(80))
def test_random_pattern_rich_circuits(seed, monkeypatch):
    monkeypatch.setenv("HQ_PEEPHOLE_X", str(seed % 2))      # odd seeds also run the opt-in x-diag-x pass
    """Gate soup over 6 qubits drawn from exactly the alphabet the patterns are made of, so matches, near misses and nested
    cases all occur; every circuit is checked against the oracle."""
    rng = random.Random(seed)
    n = 10                                   # 10 qubits = smallest tile; only the low 6 get the pattern alphabet
    lines = list(PREP)
    def soup():
        q = rng.randrange(6)
        return rng.choice([f"t q[{q}];", f"h q[{q}];", f"rz(0.7) q[{q}];", f"cx q[{q}],q[{(q + 1 + rng.randrange(5)) % 6}];", f"x q[{q}];",
                           f"cz q[{q}],q[{(q + 1 + rng.randrange(5)) % 6}];"])

    for _ in range(70):
        k = rng.random()
        a, b = rng.sample(range(6), 2)
        if k < 0.15:      # a would-be pattern with, half of the time, a random gate dropped into it (match or near miss)
            mid = [soup()] if rng.random() < 0.5 else []
            lines += [f"cx q[{a}],q[{b}];"] + mid + [f"rz({rng.uniform(0.1, 3.0):.6f}) q[{b}];"] + ([soup()] if rng.random() < 0.3 else []) + [f"cx q[{a}],q[{b}];"]
        elif k < 0.25:
            mid = [soup()] if rng.random() < 0.5 else []
            lines += [f"h q[{b}];", f"cx q[{a}],q[{b}];"] + mid + [f"cx q[{(a + 1) % 6 if (a + 1) % 6 != b else (a + 2) % 6}],q[{b}];", f"h q[{b}];"]
        elif k < 0.33:   # x ... x sandwiches around diagonal (or, as a near miss, other) gates
            lines += [f"x q[{a}];", rng.choice([f"cz q[{a}],q[{b}];", f"cu1(0.4) q[{b}],q[{a}];", f"rz(1.3) q[{a}];", soup()])]
            if rng.random() < 0.5:
                lines.append(soup())
            lines.append(f"x q[{a}];")
        elif k < 0.45:
            lines.append(f"cx q[{a}],q[{b}];")
        elif k < 0.60:
            lines.append(f"rz({rng.uniform(0.1, 3.0):.6f}) q[{a}];")
        elif k < 0.72:
            lines.append(f"h q[{a}];")
        elif k < 0.80:
            lines.append(f"t q[{a}];")
        elif k < 0.86:
            lines.append(f"z q[{a}];")
        elif k < 0.92:
            lines.append(f"cz q[{a}],q[{b}];")
        else:
            lines.append(f"ry({rng.uniform(0.1, 3.0):.6f}) q[{a}];")
    got, want, _, _ = run(qasm(n, lines))
    assert np.max(np.abs(got - want)) < 1e-12


def test_merge_pass_multiplies_adjacent_single_qubit_gates_when_cheaper():
    """u3 ; u3 -> one gate, h ; h -> nothing, t ; t -> one phase; rx(pi/2) ; t stays (a butterfly and a phase cost 3 FP64
    instructions per amplitude, their product 6); a cx in between blocks the merge.  Amplitudes as the oracle's."""
    lines = PREP + ["u3(0.3,0.1,0.2) q[0];", "u3(1.3,0.4,0.7) q[0];",        # -> 1
                    "h q[1];", "h q[1];",                                    # -> 0
                    "t q[2];", "t q[2];",                                    # -> 1
                    "rx(pi/2) q[3];", "t q[3];",                             # stays 2
                    "u3(0.3,0.1,0.2) q[5];", "cx q[5],q[6];", "u3(1.3,0.4,0.7) q[5];",   # stays 3
                    "rz(0.4) q[7];", "rz(0.9) q[7];", "rz(1.9) q[7];"]       # -> 1
    got, want, info, ngates = run(qasm(10, lines))
    assert np.max(np.abs(got - want)) < 1e-12
    # PREP's own gates merge with what follows them on qubits 0 (h;u3;u3 -> 1), 1 (ry;h;h -> ry), 2 and 7 (rx ; phases: no)
    assert info["gates"] < ngates - 5


def test_merge_pass_on_quantum_volume_matches_oracle():
    text = C.quantum_volume(12, depth=6, seed=3)
    got, want, info, ngates = run(text)
    assert np.max(np.abs(got - want)) < 1e-12
    assert info["gates"] < ngates
