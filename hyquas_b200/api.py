"""Python mirror of the reference's user API (Circuit / Gate factories, src/circuit.h:20-34, src/gate.h:30-57)
on top of the C-ABI.  Every call goes to libhyquas_b200.so; nothing here computes amplitudes.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np

from ._lib import HyquasError, check, lib

# enum class GateType order (src/gate.h:7-9)
GATE_TYPES = ["CCX", "CNOT", "CY", "CZ", "CRX", "CRY", "CU1", "CRZ", "U1", "U2", "U3", "H", "X", "Y", "Z", "S", "SDG",
              "T", "TDG", "RX", "RY", "RZ"]
_TYPE_ID = {name: i for i, name in enumerate(GATE_TYPES)}

_initialised = False


def init() -> None:
    """MyGlobalVars::init(): bind this process to its GPU (LOCAL_RANK) and, under WORLD_SIZE>1, set up NCCL."""
    global _initialised
    if not _initialised:
        check(lib.hq_runtime_init())
        _initialised = True


def init_host_only(world_size: int = 1, rank: int = 0) -> None:
    """Partitioner / planner only (no GPU is touched).  run() is unavailable in this mode."""
    check(lib.hq_runtime_init_host_only(world_size, rank))


class Circuit:
    def __init__(self, num_qubits: Optional[int] = None, _handle: Optional[ctypes.c_void_p] = None):
        if _handle is None:
            _handle = ctypes.c_void_p()
            check(lib.hq_circuit_create(int(num_qubits), ctypes.byref(_handle)))
        self._h = _handle

    @classmethod
    def from_qasm(cls, text: str) -> "Circuit":
        h = ctypes.c_void_p()
        rc = lib.hq_circuit_from_qasm(text.encode(), ctypes.byref(h))
        if rc != 0:
            raise HyquasError(lib.hq_circuit_last_error().decode())
        return cls(_handle=h)

    @property
    def num_qubits(self) -> int:
        return lib.hq_circuit_num_qubits(self._h)

    @property
    def num_gates(self) -> int:
        return lib.hq_circuit_num_gates(self._h)

    def add_gate(self, name: str, *qubits: int, params: Sequence[float] = ()) -> None:
        """add_gate('CNOT', c, t) / add_gate('CCX', c1, c2, t) / add_gate('U3', t, params=(th, ph, la)):
        operands in the reference's factory order (controls first, target last)."""
        q = list(qubits)
        t = q[-1]
        c1 = q[0] if len(q) >= 2 else -1
        c2 = q[1] if len(q) == 3 else -1
        arr = (ctypes.c_double * max(1, len(params)))(*params)
        rc = lib.hq_circuit_add_gate(self._h, _TYPE_ID[name.upper()], c2, c1, t, arr, len(params))
        if rc != 0:
            raise HyquasError(lib.hq_circuit_last_error().decode())

    def compile(self) -> None:
        check(lib.hq_circuit_compile(self._h))

    def plan_only(self):
        st, gr, sw = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        check(lib.hq_circuit_plan_only(self._h, st, gr, sw))
        return {"stages": st.value, "groups": gr.value, "swapped_bits": sw.value}

    def schedule_info(self):
        st, gr, gg = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        check(lib.hq_circuit_schedule_info(self._h, st, gr, gg))
        return {"stages": st.value, "groups": gr.value, "gates": gg.value}

    def groups(self):
        """Per gate group, in execution order: backend ('tile' | 'dense'), gates, evaluator's predicted ms, launches."""
        out = []
        for i in range(self.schedule_info()["groups"]):
            b, g, l, nb = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
            ms = ctypes.c_double()
            check(lib.hq_circuit_group_info(self._h, i, b, g, ms, l, nb))
            out.append({"backend": "dense" if b.value == 2 else "tile", "gates": g.value, "predicted_ms": ms.value,
                        "launches": l.value, "blocks": nb.value})
        return out

    def run(self, copy_back: bool = False, destroy: bool = False):
        """Circuit::run -> (wall microseconds of the execution phase, CUDA-event milliseconds)."""
        us, ms = ctypes.c_int(), ctypes.c_double()
        check(lib.hq_circuit_run(self._h, int(copy_back), int(destroy), us, ms))
        return us.value, ms.value

    def prepare_state(self) -> None:
        check(lib.hq_circuit_prepare_state(self._h))

    def execute(self, per_group: bool = False):
        """The timed phase of run() on the resident state -> (wall us, device ms[, per-group ms list])."""
        us, ms, n = ctypes.c_int(), ctypes.c_double(), ctypes.c_int()
        if per_group:
            buf = (ctypes.c_float * 4096)()
            check(lib.hq_circuit_execute(self._h, us, ms, buf, 4096, n))
            return us.value, ms.value, list(buf[:n.value])
        check(lib.hq_circuit_execute(self._h, us, ms, None, 0, n))
        return us.value, ms.value

    def release_state(self) -> None:
        """Free the resident state vector; the compiled schedule stays (prepare_state() allocates again)."""
        check(lib.hq_circuit_release_state(self._h))

    def norm2(self) -> float:
        v = ctypes.c_double()
        check(lib.hq_circuit_norm2(self._h, v))
        return v.value

    def measure(self, qubit: int) -> float:
        """Probability that the logical qubit reads 0 (every rank must call it when there are several)."""
        v = ctypes.c_double()
        check(lib.hq_circuit_measure(self._h, int(qubit), v))
        return v.value

    def io_bytes(self):
        a, b = ctypes.c_size_t(), ctypes.c_size_t()
        check(lib.hq_circuit_io_bytes(self._h, a, b))
        return a.value, b.value

    def dump(self) -> str:
        """The text printState() prints: first 128 amplitudes + every |a|^2 > 0.001 beyond."""
        need = ctypes.c_size_t()
        check(lib.hq_circuit_dump(self._h, None, 0, need))
        buf = ctypes.create_string_buffer(need.value)
        check(lib.hq_circuit_dump(self._h, buf, need.value, need))
        return buf.value.decode()

    def amplitudes(self) -> np.ndarray:
        out = np.empty(1 << self.num_qubits, dtype=np.complex128)
        rc = lib.hq_circuit_amplitudes(self._h, out.ctypes.data)
        if rc != 0:
            raise HyquasError(lib.hq_circuit_last_error().decode())
        return out

    def amp_at(self, idx: int) -> complex:
        """Circuit::ampAt: one amplitude by logical index (every rank must call it when there are several)."""
        out = (ctypes.c_double * 2)()
        rc = lib.hq_circuit_amp_at(self._h, int(idx), out)
        if rc != 0:
            raise HyquasError(lib.hq_circuit_last_error().decode())
        return complex(out[0], out[1])

    def local_shard(self, world: int) -> np.ndarray:
        """This process' shard (physical order); needs run(destroy=False)."""
        g = world.bit_length() - 1
        out = np.empty(1 << (self.num_qubits - g), dtype=np.complex128)
        check(lib.hq_circuit_local_shard(self._h, out.ctypes.data))
        return out

    def final_layout(self):
        pos = (ctypes.c_int * self.num_qubits)()
        check(lib.hq_circuit_final_layout(self._h, pos))
        return list(pos)

    def close(self) -> None:
        if self._h:
            lib.hq_circuit_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def logger_flush() -> str:
    buf = ctypes.create_string_buffer(1 << 16)
    lib.hq_circuit_logger_flush(buf, len(buf))
    return buf.value.decode()
