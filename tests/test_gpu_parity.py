"""Parity of the CUDA path (through the C-ABI) against the oracle, the reference's golden vectors and
size-independent properties.  Tolerance: max |delta amp| <= 1e-10 (FP64), BASELINE.json north_star."""
import ctypes
import os
import random

import numpy as np
import pytest

from hyquas_b200 import circuits as C
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _pack(gates):
    from hyquas_b200._lib import HqGate
    arr = (HqGate * max(1, len(gates)))()
    for i, g in enumerate(gates):
        arr[i].type, arr[i].target, arr[i].control, arr[i].control2 = 0, g.target, g.control, g.control2
        m = np.asarray(g.mat).reshape(4)
        for j in range(4):
            arr[i].mat[2 * j], arr[i].mat[2 * j + 1] = m[j].real, m[j].imag
    return arr


def _random_state(n, seed):
    r = np.random.default_rng(seed)
    s = (r.standard_normal(1 << n) + 1j * r.standard_normal(1 << n)).astype(np.complex128)
    return s / np.linalg.norm(s)


@pytest.mark.parametrize("n,K,seed", [(12, 10, 0), (13, 11, 1), (14, 12, 2), (18, 12, 3), (20, 11, 4), (20, 12, 5), (12, 12, 6)])
def test_group_kernel_vs_oracle(gpu_runtime, n, K, seed):
    """hq_group_apply on a random state and a random tile mask == gate-by-gate oracle replay."""
    from hyquas_b200._lib import check, lib
    rng = random.Random(seed)
    rest = list(range(3, n))
    rng.shuffle(rest)
    tile = [0, 1, 2] + sorted(rest[:K - 3])
    mask = sum(1 << b for b in tile)
    _, gates = O.parse_qasm(C.random_circuit(n, 300, seed))
    keep = [g for g in gates if (g.mat[0, 1] == 0 and g.mat[1, 0] == 0) or g.target in tile]
    st = _random_state(n, seed)
    want = st.copy()
    O.apply(want, n, keep)
    dev = ctypes.c_void_p()
    check(lib.hq_state_alloc(n, ctypes.byref(dev)))
    check(lib.hq_state_upload(dev, n, 0, 1 << n, st.ctypes.data))
    check(lib.hq_group_apply(dev, n, mask, _pack(keep), len(keep)))
    got = np.empty_like(st)
    check(lib.hq_state_download(dev, n, 0, 1 << n, got.ctypes.data))
    check(lib.hq_state_free(dev))
    assert np.max(np.abs(got - want)) <= 1e-13


def test_group_kernel_empty_and_scalar(gpu_runtime):
    from hyquas_b200._lib import check, lib
    n = 13
    st = _random_state(n, 11)
    dev = ctypes.c_void_p()
    check(lib.hq_state_alloc(n, ctypes.byref(dev)))
    check(lib.hq_state_upload(dev, n, 0, 1 << n, st.ctypes.data))
    check(lib.hq_group_apply(dev, n, 0xFFF, None, 0))                       # empty group: identity sweep
    scal = O.OGate("id", -1, mat=np.array([[0.6 + 0.8j, 0], [0, 0.6 + 0.8j]]))
    check(lib.hq_group_apply(dev, n, 0x17FF, _pack([scal]), 1))             # scalar on a ragged mask
    got = np.empty_like(st)
    check(lib.hq_state_download(dev, n, 0, 1 << n, got.ctypes.data))
    check(lib.hq_state_free(dev))
    assert np.max(np.abs(got - st * (0.6 + 0.8j))) <= 1e-15


def _run(api, text, keep_state=True):
    c = api.Circuit.from_qasm(text)
    c.compile()
    c.run(copy_back=False, destroy=not keep_state)
    return c


@pytest.mark.parametrize("name", ["qft_20", "bv_20", "hidden_shift_20", "supremacy_20", "quantum_volume_18", "qaoa_20",
                                  "adder_20", "basis_change_18", "supremacy_24", "qaoa_24"])
def test_circuit_vs_oracle(gpu_runtime, name):
    """Whole pipeline (parse -> partition -> lower -> kernels) vs the oracle on all 2^n amplitudes + dump text."""
    text = C.generate(name)
    c = _run(gpu_runtime, text)
    got = c.amplitudes()
    n, gates = O.parse_qasm(text)
    want = O.simulate(n, gates)
    assert np.max(np.abs(got - want)) <= TOL
    ok, err = O.compare_dumps(O.dump_state(want, n), c.dump())
    assert ok, err
    assert abs(c.norm2() - 1.0) <= 1e-10
    c.close()


@pytest.mark.parametrize("n,seed", [(10, 1), (11, 2), (15, 3), (19, 4), (22, 5)])
def test_random_circuits_all_gate_types(gpu_runtime, n, seed):
    names = ["h", "x", "y", "z", "s", "sdg", "t", "tdg", "rx", "ry", "rz", "u1", "u3", "cx", "cy", "cz", "crx", "cry",
             "crz", "cu1", "ccx"]
    text = C.random_circuit(n, 600, seed, names=names)
    c = _run(gpu_runtime, text)
    _, gates = O.parse_qasm(text)
    assert np.max(np.abs(c.amplitudes() - O.simulate(n, gates))) <= TOL
    c.close()


@pytest.mark.parametrize("name", ["qft_28", "bv_28", "hidden_shift_28"])
def test_reference_goldens_byte_exact(gpu_runtime, golden_dir, name):
    """The reference's own tests/output/*.log (configs[0] of BASELINE.json): identical text, not just within 1e-10."""
    path = os.path.join(golden_dir, name + ".qasm")
    text = open(path).read() if os.path.exists(path) else C.hidden_shift(28)
    c = _run(gpu_runtime, text, keep_state=False)
    assert c.dump() == open(os.path.join(golden_dir, name + ".log")).read()
    c.close()


@pytest.mark.parametrize("name,backend", [("supremacy_26", "b1"), ("quantum_volume_24", "b3"), ("qaoa_26", "b1"),
                                          ("basis_change_24", "b3"), ("supremacy_28", "b3"),
                                          ("supremacy_30", "b1"), ("quantum_volume_30", "b3")])   # BASELINE sizes (16 GiB state)
def test_same_dump_as_reference_build(gpu_runtime, tmp_path, name, backend):
    """Families whose upstream goldens are lost: our printState text vs the text printed by the reference's
    OWN binary (oracle/_ref/hyquas_ref_b1 = OShareMem build, b3 = TransMM build, compiled from /root/reference by
    oracle/Makefile) on the same circuit on this GPU, under scripts/compare.py's rule with a hard 1e-10 threshold."""
    import re
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "hyquas_ref_" + backend)
    if not os.path.exists(exe):
        pytest.skip("reference build not present (oracle/Makefile ref needs /root/reference)")
    text = C.generate(name)
    qasm = tmp_path / (name + ".qasm")
    qasm.write_text(text)
    # one GPU for the reference too (it drives every visible GPU otherwise, src/utils.cpp:17-60)
    r = subprocess.run([exe, str(qasm)], capture_output=True, text=True, timeout=900, env=dict(os.environ, CUDA_VISIBLE_DEVICES="0"))
    assert r.returncode == 0, r.stderr[-500:]
    ref_dump = "".join(l + "\n" for l in r.stdout.splitlines() if re.match(r"^\d+ \d\.\d+: ", l))
    assert ref_dump.count("\n") >= 128
    c = _run(gpu_runtime, text)
    ok, err = O.compare_dumps(ref_dump, c.dump())
    assert ok, err
    c.close()


def _dropin(name):
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "dropin_" + name)
    if not os.path.exists(exe):
        pytest.skip("drop-in build not present (oracle/Makefile dropin needs /root/reference)")
    return exe


@pytest.mark.parametrize("name", ["qft_28", "bv_28"])
def test_reference_main_on_our_library(gpu_runtime, golden_dir, name):
    """Source-level drop-in: the reference's OWN main.cpp, unmodified, compiled against this repo's host
    headers and linked with libhyquas_b200.so (oracle/Makefile `dropin`), run on the reference's golden circuits: the
    amplitude dump it prints must be the golden text."""
    import re
    import subprocess
    exe = _dropin("main")
    # one GPU here (on a multi-GPU box the library's launcher mode would drive all of them: tests/test_gpu_multi.py covers that)
    r = subprocess.run([exe, os.path.join(golden_dir, name + ".qasm")], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, HQ_NUM_GPUS="1"))
    assert r.returncode == 0, r.stderr[-500:]
    dump = "".join(l + "\n" for l in r.stdout.splitlines() if re.match(r"^\d+ \d\.\d+: ", l))
    assert dump == open(os.path.join(golden_dir, name + ".log")).read()
    assert re.search(r"Logger: Time Cost: \d+ us", r.stdout)


def test_reference_microbenchmark_on_our_library(gpu_runtime):
    """micro-benchmark/two-group-h.cpp of the reference, unmodified, on this library: Circuit / Gate / Logger API used the way
    the reference's own tools use it."""
    import subprocess
    exe = _dropin("two-group-h")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=dict(os.environ, HQ_NUM_GPUS="1"))
    assert r.returncode == 0, (r.stdout[-300:], r.stderr[-500:])


@pytest.mark.parametrize("name", ["supremacy_30", "qaoa_30"])
def test_full_size_round_trip(gpu_runtime, name):
    """BASELINE size (30 qubits, 16 GiB): U^dagger U |0> = |0>, norm preserved -- properties that need no oracle."""
    text = C.generate(name)
    inv = C.inverse(text)
    both = text + "".join(l + "\n" for l in inv.split("\n")[3:] if l.strip())
    c = _run(gpu_runtime, both)
    assert abs(c.norm2() - 1.0) <= 1e-10
    lines = c.dump().splitlines()
    assert len(lines) == 128                      # nothing above the 0.001 threshold beyond index 127
    idx, amp = O.parse_dump(c.dump())
    assert abs(amp[0] - 1.0) <= TOL and np.max(np.abs(amp[1:])) <= TOL
    c.close()


def test_full_size_norm_after_circuit(gpu_runtime):
    c = _run(gpu_runtime, C.generate("supremacy_30"))
    assert abs(c.norm2() - 1.0) <= 1e-10
    c.close()


def test_measure_matches_numpy(gpu_runtime):
    """hq_state_measure / Circuit.measure (kernelMeasure, src/kernelSimple.cu:482-516): P(bit = 0) for every qubit of a random
    state and after a circuit whose final layout is the identity or not; 1e-12 absolute."""
    from hyquas_b200._lib import check, lib
    n = 18
    st = _random_state(n, 21)
    dev = ctypes.c_void_p()
    check(lib.hq_state_alloc(n, ctypes.byref(dev)))
    check(lib.hq_state_upload(dev, n, 0, 1 << n, st.ctypes.data))
    idx = np.arange(1 << n)
    for t in range(n):
        p0 = ctypes.c_double()
        check(lib.hq_state_measure(dev, n, t, ctypes.byref(p0)))
        assert abs(p0.value - float(np.sum(np.abs(st[((idx >> t) & 1) == 0]) ** 2))) <= 1e-12
    check(lib.hq_state_free(dev))
    text = C.generate("qaoa_20")
    c = _run(gpu_runtime, text)
    _, gates = O.parse_qasm(text)
    want = O.simulate(20, gates)
    idx = np.arange(1 << 20)
    for q in (0, 7, 19):
        assert abs(c.measure(q) - float(np.sum(np.abs(want[((idx >> q) & 1) == 0]) ** 2))) <= 1e-12
    c.close()


def test_dump_scan_and_fetch(gpu_runtime):
    from hyquas_b200._lib import check, lib
    n = 16
    st = np.zeros(1 << n, dtype=np.complex128)
    st[5] = 0.6
    st[40000] = 0.8j
    dev = ctypes.c_void_p()
    check(lib.hq_state_alloc(n, ctypes.byref(dev)))
    check(lib.hq_state_upload(dev, n, 0, 1 << n, st.ctypes.data))
    idx = (ctypes.c_int64 * 16)()
    amp = (ctypes.c_double * 32)()
    found = ctypes.c_int64()
    check(lib.hq_dump_scan(dev, n, 0.001, idx, amp, 16, ctypes.byref(found)))
    assert found.value == 2 and list(idx[:2]) == [5, 40000]
    assert (amp[0], amp[1], amp[2], amp[3]) == (0.6, 0.0, 0.0, 0.8)
    one = (ctypes.c_double * 2)()
    check(lib.hq_amp_fetch(dev, 40000, one))
    assert (one[0], one[1]) == (0.0, 0.8)
    check(lib.hq_state_free(dev))


# ---- fused dense-matrix (TransMM-class) kernel -----------------------------------------------------------------------
def _apply_dense_np(state, n, qubits, U):
    m = len(qubits)
    psi = state.reshape([2] * n)
    axes = [n - 1 - q for q in reversed(qubits)]
    psi = np.moveaxis(psi, axes, range(m))
    shp = psi.shape
    psi = (U @ psi.reshape(1 << m, -1)).reshape(shp)
    return np.ascontiguousarray(np.moveaxis(psi, range(m), axes)).reshape(-1)


def _pack_u(mats):
    out = []
    for U in mats:
        cm = np.asarray(U).T.reshape(-1)
        out.append(np.stack([cm.real, cm.imag], axis=1).reshape(-1))
    return np.ascontiguousarray(np.concatenate(out))


@pytest.mark.parametrize("n,groups", [
    (12, [[0, 1, 2]]), (14, [[3, 7, 11]]), (16, [[5, 9, 10, 12]]), (18, [[0, 4, 8, 12, 17]]), (18, [[2, 3, 5, 7, 11, 13]]),
    (15, [[6]]), (15, [[1, 9]]), (20, [[14, 15, 16, 17, 18, 19]]), (20, [[3, 6, 9, 12]]),
    (20, [[4, 5, 6], [6, 7, 8, 9], [0, 15]]), (17, [[10, 11, 12, 13], [3, 10, 11, 12, 13]]), (13, [[4, 5, 6, 7, 8, 9], [0, 1, 2, 3, 10, 11]]),
    (22, [[0, 1, 2, 3, 4, 5]]), (22, [[16, 17, 18, 19, 20, 21], [3, 4, 5, 16]]),
])
def test_dense_kernel_vs_numpy(gpu_runtime, n, groups):
    """hq_dense_plan_launch (DMMA kernel) == numpy U @ x on the chosen qubits, random unitary, random state."""
    from hyquas_b200._lib import check, lib
    rng = np.random.default_rng(n * 31 + len(groups))
    st = _random_state(n, n)
    mats = []
    for q in groups:
        a = rng.standard_normal((1 << len(q), 1 << len(q))) + 1j * rng.standard_normal((1 << len(q), 1 << len(q)))
        mats.append(np.linalg.qr(a)[0])
    want = st.copy()
    for q, U in zip(groups, mats):
        want = _apply_dense_np(want, n, q, U)
    m_list = (ctypes.c_int * len(groups))(*[len(q) for q in groups])
    flat = [b for q in groups for b in q]
    u = _pack_u(mats)
    plan, dev = ctypes.c_void_p(), ctypes.c_void_p()
    check(lib.hq_dense_plan_create(n, len(groups), m_list, (ctypes.c_int * len(flat))(*flat), u.ctypes.data, ctypes.byref(plan)))
    check(lib.hq_state_alloc(n, ctypes.byref(dev)))
    check(lib.hq_state_upload(dev, n, 0, 1 << n, st.ctypes.data))
    check(lib.hq_dense_plan_launch(plan, dev, 0))
    got = np.empty_like(st)
    check(lib.hq_state_download(dev, n, 0, 1 << n, got.ctypes.data))
    check(lib.hq_state_free(dev))
    lib.hq_dense_plan_destroy(plan)
    assert np.max(np.abs(got - want)) <= 1e-13
