"""ctypes binding of libhyquas_b200.so (the C-ABI in include/hyquas_b200.h and hyquas_b200_circuit.h).

There is no fallback: if the shared library is missing or a symbol is absent, importing this module raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# HQ_LIB_SUFFIX selects an alternative build of the same library (kernel A/B experiments, e.g. "_r4")
LIB_PATH = os.path.join(_HERE, "libhyquas_b200" + os.environ.get("HQ_LIB_SUFFIX", "") + ".so")


class HqGate(ctypes.Structure):
    _fields_ = [("type", ctypes.c_int32), ("target", ctypes.c_int32), ("control", ctypes.c_int32),
                ("control2", ctypes.c_int32), ("mat", ctypes.c_double * 8)]


class HyquasError(RuntimeError):
    pass


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `make -C hyquas_b200/csrc` (or __graft_entry__.build()); "
            "hyquas_b200 has no CPU fallback")
    return ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)


lib = _load()

_c = ctypes
_P = _c.POINTER
_SIGS = {
    # device layer -- include/hyquas_b200.h
    "hq_last_error": (_c.c_char_p, []),
    "hq_version": (_c.c_char_p, []),
    "hq_device_count": (_c.c_int, [_P(_c.c_int)]),
    "hq_init": (_c.c_int, [_c.c_int]),
    "hq_shutdown": (_c.c_int, []),
    "hq_sync": (_c.c_int, []),
    "hq_device_info": (_c.c_int, [_c.c_char_p, _c.c_size_t, _P(_c.c_int), _P(_c.c_size_t)]),
    "hq_state_alloc": (_c.c_int, [_c.c_int, _P(_c.c_void_p)]),
    "hq_state_free": (_c.c_int, [_c.c_void_p]),
    "hq_state_init": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int]),
    "hq_state_download": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int64, _c.c_int64, _c.c_void_p]),
    "hq_state_upload": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int64, _c.c_int64, _c.c_void_p]),
    "hq_amp_fetch": (_c.c_int, [_c.c_void_p, _c.c_int64, _P(_c.c_double)]),
    "hq_dump_scan": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_double, _c.c_void_p, _c.c_void_p, _c.c_int64, _P(_c.c_int64)]),
    "hq_state_measure": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _P(_c.c_double)]),
    "hq_state_norm2": (_c.c_int, [_c.c_void_p, _c.c_int, _P(_c.c_double)]),
    "hq_group_tile_bits": (_c.c_int, []),
    "hq_group_min_run_bits": (_c.c_int, []),
    "hq_group_plan_create": (_c.c_int, [_c.c_int, _c.c_uint64, _P(HqGate), _c.c_int, _P(_c.c_void_p)]),
    "hq_group_plan_create_ex": (_c.c_int, [_c.c_int, _c.c_uint64, _c.c_uint64, _c.c_uint64, _P(HqGate), _c.c_int, _P(_c.c_void_p)]),
    "hq_group_plan_launch": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int]),
    "hq_group_plan_info": (_c.c_int, [_c.c_void_p, _P(_c.c_int), _P(_c.c_int), _P(_c.c_int), _P(_c.c_int)]),
    "hq_group_plan_table_bytes": (_c.c_int, [_c.c_void_p, _P(_c.c_int)]),
    "hq_group_plan_local_exchanges": (_c.c_int, [_c.c_void_p, _P(_c.c_int)]),
    "hq_group_plan_destroy": (_c.c_int, [_c.c_void_p]),
    "hq_group_plan_cost": (_c.c_int, [_c.c_void_p, _P(_c.c_int), _P(_c.c_double)]),
    "hq_group_plan_enable_zero_input": (_c.c_int, [_c.c_void_p]),
    "hq_group_plan_launch_from_zero": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int]),
    "hq_group_plans_warm": (_c.c_int, [_P(_c.c_void_p), _c.c_int]),
    "hq_group_plan_is_specialised": (_c.c_int, [_c.c_void_p, _P(_c.c_int)]),
    "hq_jit_available": (_c.c_int, [_P(_c.c_int)]),
    "hq_cache_dir": (_c.c_int, [_c.c_char_p, _c.c_size_t]),
    "hq_jit_stats": (_c.c_int, [_P(_c.c_int), _P(_c.c_int), _P(_c.c_int), _P(_c.c_double)]),
    "hq_group_apply": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_uint64, _P(HqGate), _c.c_int]),
    "hq_microbench_fp64": (_c.c_int, [_c.c_int, _P(_c.c_double)]),
    "hq_microbench_copy": (_c.c_int, [_c.c_void_p, _c.c_int, _P(_c.c_double)]),
    "hq_dense_plan_create": (_c.c_int, [_c.c_int, _c.c_int, _P(_c.c_int), _P(_c.c_int), _c.c_void_p, _P(_c.c_void_p)]),
    "hq_dense_plan_create_ex": (_c.c_int, [_c.c_int, _c.c_uint64, _c.c_uint64, _c.c_int, _P(_c.c_int), _P(_c.c_int), _c.c_void_p,
                                           _P(_c.c_void_p)]),
    "hq_dense_plan_launch": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_int]),
    "hq_dense_plan_info": (_c.c_int, [_c.c_void_p, _P(_c.c_int), _P(_c.c_int), _P(_c.c_int), _P(_c.c_double), _P(_c.c_int)]),
    "hq_dense_plan_destroy": (_c.c_int, [_c.c_void_p]),
    "hq_dense_apply": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _P(_c.c_int), _c.c_void_p]),
    "hq_comm_unique_id": (_c.c_int, [_c.c_void_p]),
    "hq_comm_init": (_c.c_int, [_c.c_int, _c.c_int, _c.c_void_p]),
    "hq_comm_info": (_c.c_int, [_P(_c.c_int), _P(_c.c_int)]),
    "hq_comm_destroy": (_c.c_int, []),
    "hq_comm_bcast_host": (_c.c_int, [_c.c_void_p, _c.c_size_t, _c.c_int]),
    "hq_comm_allgather_host": (_c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_size_t]),
    "hq_swap_any_position": (_c.c_int, [_P(_c.c_int)]),
    "hq_swap_attach": (_c.c_int, [_c.c_void_p]),
    "hq_swap_detach": (_c.c_int, []),
    "hq_state_bitswap": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _P(_c.c_int), _P(_c.c_int)]),
    "hq_swap_plan_create": (_c.c_int, [_c.c_int, _c.c_int, _P(_c.c_int), _P(_c.c_int), _P(_c.c_void_p)]),
    "hq_swap_plan_set_overlap": (_c.c_int, [_c.c_void_p, _c.c_int]),
    "hq_swap_begin": (_c.c_int, [_c.c_void_p, _c.c_void_p]),
    "hq_swap_wait_chunk": (_c.c_int, [_c.c_void_p, _P(_c.c_int)]),
    "hq_swap_end": (_c.c_int, [_c.c_void_p]),
    "hq_swap_plan_destroy": (_c.c_int, [_c.c_void_p]),
    "hq_timer_start": (_c.c_int, []),
    "hq_timer_stop_ms": (_c.c_int, [_P(_c.c_float)]),
    # circuit layer -- include/hyquas_b200_circuit.h
    "hq_circuit_last_error": (_c.c_char_p, []),
    "hq_runtime_init": (_c.c_int, []),
    "hq_runtime_init_host_only": (_c.c_int, [_c.c_int, _c.c_int]),
    "hq_circuit_create": (_c.c_int, [_c.c_int, _P(_c.c_void_p)]),
    "hq_circuit_from_qasm": (_c.c_int, [_c.c_char_p, _P(_c.c_void_p)]),
    "hq_circuit_add_gate": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _P(_c.c_double), _c.c_int]),
    "hq_circuit_num_qubits": (_c.c_int, [_c.c_void_p]),
    "hq_circuit_num_gates": (_c.c_int, [_c.c_void_p]),
    "hq_circuit_compile": (_c.c_int, [_c.c_void_p]),
    "hq_circuit_plan_only": (_c.c_int, [_c.c_void_p, _P(_c.c_int), _P(_c.c_int), _P(_c.c_int)]),
    "hq_circuit_run": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _P(_c.c_int), _P(_c.c_double)]),
    "hq_circuit_prepare_state": (_c.c_int, [_c.c_void_p]),
    "hq_circuit_execute": (_c.c_int, [_c.c_void_p, _P(_c.c_int), _P(_c.c_double), _P(_c.c_float), _c.c_int, _P(_c.c_int)]),
    "hq_circuit_measure": (_c.c_int, [_c.c_void_p, _c.c_int, _P(_c.c_double)]),
    "hq_circuit_release_state": (_c.c_int, [_c.c_void_p]),
    "hq_circuit_norm2": (_c.c_int, [_c.c_void_p, _P(_c.c_double)]),
    "hq_circuit_swap_alone_ms": (_c.c_int, [_c.c_void_p, _P(_c.c_double)]),
    "hq_circuit_io_bytes": (_c.c_int, [_c.c_void_p, _P(_c.c_size_t), _P(_c.c_size_t)]),
    "hq_circuit_schedule_info": (_c.c_int, [_c.c_void_p, _P(_c.c_int), _P(_c.c_int), _P(_c.c_int)]),
    "hq_circuit_group_info": (_c.c_int, [_c.c_void_p, _c.c_int, _P(_c.c_int), _P(_c.c_int), _P(_c.c_double), _P(_c.c_int), _P(_c.c_int)]),
    "hq_circuit_group_cost": (_c.c_int, [_c.c_void_p, _c.c_int, _P(_c.c_int), _P(_c.c_double)]),
    "hq_circuit_dump": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.c_size_t, _P(_c.c_size_t)]),
    "hq_circuit_amplitudes": (_c.c_int, [_c.c_void_p, _c.c_void_p]),
    "hq_circuit_amp_at": (_c.c_int, [_c.c_void_p, _c.c_longlong, _P(_c.c_double)]),
    "hq_circuit_local_shard": (_c.c_int, [_c.c_void_p, _c.c_void_p]),
    "hq_circuit_final_layout": (_c.c_int, [_c.c_void_p, _P(_c.c_int)]),
    "hq_circuit_logger_flush": (_c.c_int, [_c.c_char_p, _c.c_size_t]),
    "hq_circuit_destroy": (_c.c_int, [_c.c_void_p]),
    # test hook (device/plan_emulator.cpp) -- used by the CPU test-suite only
    "hq_debug_group_plan_emulate": (_c.c_int, [_c.c_void_p, _c.c_void_p]),
    "hq_debug_group_plan_jit_source": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_char_p, _c.c_size_t, _P(_c.c_size_t)]),
    "hq_debug_jit_compile_to_file": (_c.c_int, [_c.c_char_p, _c.c_char_p, _c.c_char_p, _c.c_size_t]),
    "hq_debug_jit_cache_probe": (_c.c_int, [_c.c_char_p, _c.c_char_p, _P(_c.c_int), _P(_c.c_int)]),
    "hq_debug_dense_plan_emulate": (_c.c_int, [_c.c_void_p, _c.c_void_p]),
    "hq_debug_schedule_check": (_c.c_int, [_c.c_void_p, _c.c_char_p, _c.c_size_t]),
    "hq_debug_num_stages": (_c.c_int, [_c.c_void_p]),
    "hq_debug_stage_swap": (_c.c_int, [_c.c_void_p, _c.c_int, _P(_c.c_int), _P(_c.c_int), _P(_c.c_int), _P(_c.c_int),
                                       _P(_c.c_int), _P(_c.c_int), _P(_c.c_int)]),
    "hq_debug_stage_emulate": (_c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_void_p]),
    "hq_debug_final_pos": (_c.c_int, [_c.c_void_p, _P(_c.c_int)]),
}

for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)      # AttributeError here == the library does not export a declared symbol
    _fn.restype = _res
    _fn.argtypes = _args


def check(rc: int) -> None:
    if rc != 0:
        msg = lib.hq_last_error().decode() or lib.hq_circuit_last_error().decode()
        raise HyquasError(f"hyquas_b200 error {rc}: {msg}")
