"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's observable behaviour for one circuit:
QASM-subset text -> gate list -> gate-by-gate FP64 replay (oracle.c) -> amplitude dump text.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under hyquas_b200/ does.

What it follows (paths relative to /root/reference):
  * input grammar + angle parsing     main.cpp:12-66 (parse_qid, parse_gate), main.cpp:68-231 (token loop)
  * gate matrices                      src/gate.cpp:9-342  (one entry per factory, cited in GATE_MATRIX)
  * replay index rule                  src/kernelSimple.cu:40-70,187-198 (in oracle.c)
  * dump format                        src/circuit.h:14-16 (ResultItem::print), src/circuit.cpp:284-309 (printState),
                                       src/utils.cpp:77-84 (zero_wrapper)
  * compare tolerance handling         scripts/compare.py:15,25,29-30

Parity pinning: oracle output == tests/golden/{qft_28,bv_28,hidden_shift_28}.log byte for byte
(tests/test_oracle_golden.py; the n=28 replays are gated behind HYQUAS_SLOW=1 because they take
minutes on 8 cores; n<=22 analytic forms of the same families run in the default CPU suite).

Deliberate deviation: `tdg` uses the matrix of src/gate.cpp:236-246 (diag(1, e^{-i pi/4})).  The
reference tags that gate with type T (gate.cpp:239), so its OShareMem kernel applies T instead
(kernelOpt.cu:350) while its own BLAS path applies the correct matrix; we follow the matrix.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


# ----------------------------------------------------------------------------------------------
# gate matrices, src/gate.cpp
# ----------------------------------------------------------------------------------------------
_R2 = 1.0 / math.sqrt(2)


def _m(a, b, c, d):
    return np.array([[a, b], [c, d]], dtype=np.complex128)


def gate_matrix(name: str, p: Sequence[float] = ()) -> np.ndarray:
    """2x2 matrix exactly as the reference's Gate factories fill `mat` (src/gate.cpp)."""
    cos, sin = math.cos, math.sin
    if name in ("x", "cx", "ccx"):           # gate.cpp:9-33,164-174
        return _m(0, 1, 1, 0)
    if name in ("y", "cy"):                   # gate.cpp:35-45,176-186
        return _m(0, -1j, 1j, 0)
    if name in ("z", "cz"):                   # gate.cpp:47-57,188-198
        return _m(1, 0, 0, -1)
    if name in ("rx", "crx"):                 # gate.cpp:59-69,248-258
        a = p[0]
        return _m(cos(a / 2.0), complex(0, -sin(a / 2.0)), complex(0, -sin(a / 2.0)), cos(a / 2.0))
    if name in ("ry", "cry"):                 # gate.cpp:71-81,260-270
        a = p[0]
        return _m(cos(a / 2.0), -sin(a / 2.0), sin(a / 2.0), cos(a / 2.0))
    if name in ("rz", "crz"):                 # gate.cpp:97-107,272-282
        a = p[0]
        return _m(complex(cos(a / 2), -sin(a / 2)), 0, 0, complex(cos(a / 2), sin(a / 2)))
    if name in ("u1", "cu1"):                 # gate.cpp:83-95,110-122
        l = p[0]
        return _m(1, 0, 0, complex(cos(l), sin(l)))
    if name == "u2":                          # gate.cpp:124-136
        phi, lam = p
        s2 = math.sqrt(2)
        return _m(1.0 / s2, complex(-cos(lam) / s2, -sin(lam) / s2),
                  complex(cos(phi) / s2, sin(phi) / s2), complex(cos(lam + phi) / s2, sin(lam + phi) / s2))
    if name == "u3":                          # gate.cpp:138-150
        th, phi, lam = p
        return _m(cos(th / 2), complex(-cos(lam) * sin(th / 2), -sin(lam) * sin(th / 2)),
                  complex(cos(phi) * sin(th / 2), sin(phi) * sin(th / 2)),
                  complex(cos(phi + lam) * cos(th / 2), sin(phi + lam) * cos(th / 2)))
    if name == "h":                           # gate.cpp:152-162
        return _m(_R2, _R2, _R2, -_R2)
    if name == "s":                           # gate.cpp:200-210
        return _m(1, 0, 0, 1j)
    if name == "sdg":                         # gate.cpp:212-222
        return _m(1, 0, 0, -1j)
    if name == "t":                           # gate.cpp:224-234
        return _m(1, 0, 0, complex(_R2, _R2))
    if name == "tdg":                         # gate.cpp:236-246 (matrix; see module docstring)
        return _m(1, 0, 0, complex(_R2, -_R2))
    if name == "id":                          # gate.cpp:284-294
        return _m(1, 0, 0, 1)
    raise ValueError(f"unknown gate {name}")


@dataclass
class OGate:
    name: str                 # qasm token (lower case)
    target: int
    control: int = -1
    control2: int = -1
    params: Tuple[float, ...] = ()
    mat: np.ndarray = field(default=None, repr=False)

    def __post_init__(self):
        if self.mat is None:
            self.mat = gate_matrix(self.name, self.params)


# ----------------------------------------------------------------------------------------------
# parser, main.cpp
# ----------------------------------------------------------------------------------------------
def _parse_qid(tok: str) -> List[int]:
    """main.cpp:12-27 -- every maximal digit run in the operand token is a qubit id."""
    out, i = [], 0
    while i < len(tok):
        if tok[i].isdigit():
            j = i
            while j < len(tok) and tok[j].isdigit():
                j += 1
            out.append(int(tok[i:j]))
            i = j
        else:
            i += 1
    return out


def _parse_param(st: str) -> float:
    """main.cpp:49-61 -- `pi*x` -> pi*x, `pi/x` -> pi/x, anything else -> stod."""
    pi = math.acos(-1)
    if st.startswith("pi*"):
        return pi * float(st[3:])
    if st.startswith("pi/"):
        return pi / float(st[3:])
    return 1.0 * float(st)


_FIXED = {"cx": 2, "ccx": 3, "cy": 2, "cz": 2, "h": 1, "x": 1, "y": 1, "z": 1, "s": 1, "sdg": 1, "t": 1, "tdg": 1}
_PARAM = {"crx": (1, 2), "cry": (1, 2), "crz": (1, 2), "cu1": (1, 2), "u1": (1, 1), "u3": (3, 1),
          "rx": (1, 1), "ry": (1, 1), "rz": (1, 1)}


def parse_qasm(text: str) -> Tuple[int, List[OGate]]:
    """Token loop of main.cpp:68-231: whitespace tokens, rest of line dropped after each statement."""
    n = -1
    gates: List[OGate] = []
    for raw in text.split("\n"):
        toks = raw.split()
        if not toks:
            continue
        head = toks[0]
        if head in ("//", "OPENQASM", "include"):     # main.cpp:77
            continue
        if head == "qreg":                              # main.cpp:78-80: "%*c%*c%*c%d" skips ' q['
            n = int(_parse_qid(toks[1])[0])
            continue
        if head in _FIXED:                              # main.cpp:81-152
            q = _parse_qid(toks[1])
            assert len(q) == _FIXED[head]
            if len(q) == 1:
                gates.append(OGate(head, q[0]))
            elif len(q) == 2:
                gates.append(OGate(head, q[1], q[0]))
            else:
                gates.append(OGate(head, q[2], q[0], q[1]))
            continue
        name = head.split("(")[0]                       # main.cpp:29-66
        if name not in _PARAM:
            raise SystemExit(f"unrecognized token {head}")   # main.cpp:218-221
        inner = head[head.index("(") + 1: head.rindex(")")]
        params = tuple(_parse_param(s) for s in inner.split(","))
        npar, nq = _PARAM[name]
        assert len(params) == npar
        q = _parse_qid(toks[1])
        assert len(q) == nq
        if nq == 1:
            gates.append(OGate(name, q[0], params=params))
        else:
            gates.append(OGate(name, q[1], q[0], params=params))
    if n < 0:
        raise SystemExit("fail to load circuit")       # main.cpp:226-229
    return n, gates


# ----------------------------------------------------------------------------------------------
# C replay library
# ----------------------------------------------------------------------------------------------
class _CGate(ctypes.Structure):
    _fields_ = [("target", ctypes.c_int32), ("control", ctypes.c_int32), ("control2", ctypes.c_int32),
                ("pad", ctypes.c_int32), ("m", ctypes.c_double * 8)]


_lib = None


def build(force: bool = False) -> str:
    """gcc -O2 -fopenmp oracle.c -> liboracle.so (recipe also in oracle/Makefile)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", _LIB_PATH, src, "-lm"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.orc_run.restype = ctypes.c_double
        L.orc_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(_CGate), ctypes.c_int]
        L.orc_init_zero_state.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.orc_scan_large.restype = ctypes.c_int64
        L.orc_scan_large.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_int64,
                                     ctypes.c_void_p, ctypes.c_int64]
        L.orc_norm2.restype = ctypes.c_double
        L.orc_norm2.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.orc_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _pack(gates: Sequence[OGate]):
    arr = (_CGate * max(1, len(gates)))()
    for k, g in enumerate(gates):
        arr[k].target, arr[k].control, arr[k].control2 = g.target, g.control, g.control2
        m = np.asarray(g.mat, dtype=np.complex128).reshape(4)
        for j in range(4):
            arr[k].m[2 * j] = m[j].real
            arr[k].m[2 * j + 1] = m[j].imag
    return arr


def zero_state(n: int) -> np.ndarray:
    s = np.empty(1 << n, dtype=np.complex128)
    lib().orc_init_zero_state(s.ctypes.data, n)
    return s


def apply(state: np.ndarray, n: int, gates: Sequence[OGate]) -> float:
    """In-place replay; returns seconds spent in the gate loop (what 'Time Cost' would cover)."""
    assert state.dtype == np.complex128 and state.flags["C_CONTIGUOUS"] and state.size == 1 << n
    return lib().orc_run(state.ctypes.data, n, _pack(gates), len(gates))


def simulate(n: int, gates: Sequence[OGate]) -> np.ndarray:
    s = zero_state(n)
    apply(s, n, gates)
    return s


def simulate_qasm(text: str) -> Tuple[int, np.ndarray]:
    n, gates = parse_qasm(text)
    return n, simulate(n, gates)


def simulate_numpy(n: int, gates: Sequence[OGate]) -> np.ndarray:
    """Independent pure-numpy replay (small n only) used to cross-check oracle.c itself."""
    s = np.zeros(1 << n, dtype=np.complex128)
    s[0] = 1.0
    idx = np.arange(1 << n)
    for g in gates:
        t = g.target
        sel = (idx >> t) & 1 == 0
        if g.control >= 0:
            sel &= (idx >> g.control) & 1 == 1
        if g.control2 >= 0:
            sel &= (idx >> g.control2) & 1 == 1
        lo = idx[sel]
        hi = lo | (1 << t)
        a, b = s[lo].copy(), s[hi].copy()
        s[lo] = g.mat[0, 0] * a + g.mat[0, 1] * b
        s[hi] = g.mat[1, 0] * a + g.mat[1, 1] * b
    return s


# ----------------------------------------------------------------------------------------------
# dump, src/circuit.cpp:284-309
# ----------------------------------------------------------------------------------------------
def _zw(x: float) -> float:
    """zero_wrapper, src/utils.cpp:77-84."""
    return 0.0 if -1e-14 < x < 1e-14 else x


def format_item(idx: int, amp: complex) -> str:
    """ResultItem::print, src/circuit.h:14-16."""
    return "%d %.12f: %.12f %.12f\n" % (idx, amp.real * amp.real + amp.imag * amp.imag, _zw(amp.real), _zw(amp.imag))


def dump_state(state: np.ndarray, n: int) -> str:
    """printState (non-MPI, SHOW_SCHEDULE off): first 128 logical amplitudes, then every amplitude with
    |a|^2 > 0.001 and index >= 128, ascending."""
    out = [format_item(i, complex(state[i])) for i in range(128)]
    if state.size > 128:
        p = state.real[128:] ** 2 + state.imag[128:] ** 2 if state.size <= (1 << 24) else None
        if p is not None:
            big = np.nonzero(p > 0.001)[0] + 128
        else:
            cap = 2048
            buf = np.empty(cap, dtype=np.int64)
            cnt = lib().orc_scan_large(state.ctypes.data, n, 0.001, 128, buf.ctypes.data, cap)
            assert cnt <= cap
            big = buf[:cnt]
        out += [format_item(int(i), complex(state[i])) for i in big]
    return "".join(out)


def parse_dump(text: str) -> Tuple[np.ndarray, np.ndarray]:
    """-> (indices, complex amplitudes); skips Logger lines like scripts/compare.py:19-21."""
    idx, amp = [], []
    for line in text.splitlines():
        if not line.strip() or line.startswith("Logger"):
            continue
        f = line.split()
        idx.append(int(f[0]))
        amp.append(complex(float(f[2]), float(f[3])))
    return np.array(idx, dtype=np.int64), np.array(amp, dtype=np.complex128)


def compare_dumps(std: str, mine: str, tol: float = 1e-10) -> Tuple[bool, float]:
    """scripts/compare.py:8-33 with a hard threshold: clamp |x|<1e-10 to 0, max abs error <= tol."""
    i0, a0 = parse_dump(std)
    i1, a1 = parse_dump(mine)
    if i0.shape != i1.shape or not np.array_equal(i0, i1):
        return False, float("inf")
    v0 = np.stack([a0.real, a0.imag], 1)
    v1 = np.stack([a1.real, a1.imag], 1)
    v0[np.abs(v0) < 1e-10] = 0
    v1[np.abs(v1) < 1e-10] = 0
    err = float(np.max(np.abs(v0 - v1))) if v0.size else 0.0
    return err <= tol, err
