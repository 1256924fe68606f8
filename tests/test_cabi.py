"""The C-ABI library loads and exports every symbol include/*.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = []
    for hdr in ("hyquas_b200.h", "hyquas_b200_circuit.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"\b(hq_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_every_declared_symbol_is_exported():
    from hyquas_b200._lib import LIB_PATH
    lib = ctypes.CDLL(LIB_PATH)
    declared = _declared()
    assert len(declared) >= 40
    missing = [n for n in declared if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_table_covers_header():
    from hyquas_b200 import _lib
    declared = set(_declared())
    bound = set(_lib._SIGS)
    assert declared <= bound, sorted(declared - bound)


def test_errors_are_reported_not_swallowed():
    from hyquas_b200._lib import lib
    h = ctypes.c_void_p()
    assert lib.hq_circuit_from_qasm(b"qreg q[4];\nswap q[0],q[1];\n", ctypes.byref(h)) != 0
    assert b"unrecognized token" in lib.hq_circuit_last_error()
    assert lib.hq_circuit_from_qasm(b"h q[0];\n", ctypes.byref(h)) != 0
    plan = ctypes.c_void_p()
    # tile mask without the low bits / wrong popcount
    assert lib.hq_group_plan_create(14, 0b111111111100, None, 0, ctypes.byref(plan)) != 0
    assert lib.hq_group_plan_create(14, 0xFFFF, None, 0, ctypes.byref(plan)) != 0


def test_run_without_gpu_fails_loudly():
    """No CPU fallback: launching a plan with no bound GPU is an error, never a silent host computation."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hyquas_b200._lib import lib
    plan = ctypes.c_void_p()
    assert lib.hq_group_plan_create(12, 0x3FF, None, 0, ctypes.byref(plan)) == 0
    buf = (ctypes.c_double * 8)()
    assert lib.hq_group_plan_launch(plan, buf, 0) != 0
    assert b"hq_init" in lib.hq_last_error()
    lib.hq_group_plan_destroy(plan)


def test_reference_sources_compile_against_our_headers():
    """main.cpp and micro-benchmark/*.cpp of the reference "must compile unchanged" against this repo's host
    headers.  Only where the reference tree is mounted (this container); the GPU box runs the resulting binaries instead
    (tests/test_gpu_parity.py::test_reference_main_on_our_library)."""
    import os
    import subprocess
    import pytest
    if not os.path.isdir("/root/reference/micro-benchmark"):
        pytest.skip("reference sources not mounted")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["make", "-C", os.path.join(root, "oracle"), "dropin"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for name in ("main", "local-single", "local-ctr", "two-group-h", "bench-blas"):
        assert os.path.exists(os.path.join(root, "oracle", "_ref", "dropin_" + name))
