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


def test_add_gate_rejects_bad_operands():
    """Operands must be in range and distinct, controlled types need their controls: an error code, not a dead process."""
    from hyquas_b200 import api
    from hyquas_b200._lib import HyquasError
    c = api.Circuit(5)
    c.add_gate("CNOT", 0, 1)
    for bad in [("CNOT", (1, 1)), ("CCX", (0, 0, 2)), ("CCX", (0, 1, 1)), ("H", (7,)), ("CZ", (2, 9))]:
        with pytest.raises(HyquasError):
            c.add_gate(bad[0], *bad[1])
    import ctypes
    from hyquas_b200._lib import lib
    arr = (ctypes.c_double * 1)(0.0)
    assert lib.hq_circuit_add_gate(c._h, 1, -1, -1, 2, arr, 0) != 0     # CNOT without a control
    assert lib.hq_circuit_add_gate(c._h, 11, -1, 3, 2, arr, 0) != 0     # H with a control
    assert c.num_gates == 1
    with pytest.raises(HyquasError):
        api.Circuit.from_qasm('OPENQASM 2.0;\ninclude "qelib1.inc";\nqreg q[4];\ncx q[1],q[1];\n')


def test_compile_refuses_too_few_local_qubits_with_an_error_code():
    """12 qubits on 8 GPUs leave 9 local qubits: the tile kernel needs 10; compile() must say so, not exit."""
    import subprocess
    import sys
    code = ("from hyquas_b200 import api, circuits\n"
            "from hyquas_b200._lib import HyquasError\n"
            "api.init_host_only(8, 0)\n"
            "c = api.Circuit.from_qasm(circuits.random_circuit(12, 60, 1))\n"
            "try:\n    c.compile()\nexcept HyquasError as e:\n    print('refused:', e)\n"
            "try:\n    c.plan_only()\nexcept HyquasError as e:\n    print('refused plan:', e)\n"
            "api.init_host_only(8, 0)\n"
            "c2 = api.Circuit.from_qasm(circuits.random_circuit(13, 60, 1)); print(c2.plan_only())\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-800:]
    assert "refused:" in r.stdout and "refused plan:" in r.stdout and "at least 10 are needed" in r.stdout
    assert "'stages'" in r.stdout


def _param_file(path, h_us):
    """A {L}qubits.out in the reference's layout (evaluator-preprocess/process.cpp:167-178) with hq_preprocess's marker."""
    single = [100000] * 14
    single[3] = h_us                      # H
    lines = ["1"] + [f"{v} " for v in single] + [""] + [f"{v} " for v in [90000] * 7]
    lines += [f"{1 << m} {0.5 * (m + 1):f}" for m in range(10)] + ["", "0.000000", "#hq_preprocess test"]
    open(path, "w").write("\n".join(lines) + "\n")


def test_evaluator_reads_reference_layout_parameter_files(tmp_path):
    """Evaluator::loadParam reads {L}qubits.out files in the reference's layout (what hq_preprocess writes): the predicted time of
    an H-heavy group follows the H entry of the file; a file without the marker on the default path would be ignored."""
    import subprocess
    import sys
    code = ("from hyquas_b200 import api\n"
            "api.init_host_only(1, 0)\n"
            "c = api.Circuit(20)\n"
            "for i in range(200): c.add_gate('H', 5 + i % 4); c.add_gate('T', 5 + i % 4)\n"
            "c.compile(); print('PRED', sum(g['predicted_ms'] for g in c.groups()))\n")
    preds = []
    for h_us in (20000, 2000000):
        d = tmp_path / f"p{h_us}"
        d.mkdir()
        _param_file(str(d / "20qubits.out"), h_us)
        env = dict(os.environ, HYQUAS_PARAM_DIR=str(d), HQ_PEEPHOLE="0", HQ_BACKEND="group")
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
        assert r.returncode == 0, r.stderr[-500:]
        preds.append(float(re.search(r"PRED ([0-9.eE+-]+)", r.stdout).group(1)))
    assert preds[1] > 10 * preds[0] > 0


@pytest.mark.gpu
def test_hq_preprocess_writes_the_reference_layout(tmp_path):
    """hq_preprocess (the reference's `process` tool for this library): one file per L with param_type, 14 + 7 integer times,
    ten `K ms` lines, one transpose line -- exactly what src/evaluator.cpp:60-103 parses -- plus the marker line."""
    import subprocess
    exe = os.path.join(ROOT, "hyquas_b200", "hq_preprocess")
    r = subprocess.run([exe, "20"], capture_output=True, text=True, timeout=900, env=dict(os.environ, HYQUAS_PARAM_DIR=str(tmp_path)))
    assert r.returncode == 0, (r.stdout[-300:], r.stderr[-500:])
    tok = open(tmp_path / "20qubits.out").read().split("#hq_preprocess")[0].split()
    assert tok[0] == "1" and len(tok) == 1 + 14 + 7 + 20 + 1
    times = [int(t) for t in tok[1:22]]
    assert all(t > 0 for t in times)
    assert [int(tok[22 + 2 * i]) for i in range(10)] == [1 << i for i in range(10)]
    assert all(float(tok[23 + 2 * i]) > 0 for i in range(10))
