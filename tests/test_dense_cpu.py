"""Dense (TransMM-class) plan on CPU: the planner's tables (swizzle, index orders, fragment-ordered U, padding) are
interpreted by the plan emulator (test hook) and compared with a direct numpy application of U on the chosen qubits."""
import ctypes

import numpy as np
import pytest

from hyquas_b200._lib import check, lib


def random_unitary(k, rng):
    a = rng.standard_normal((k, k)) + 1j * rng.standard_normal((k, k))
    q, _ = np.linalg.qr(a)
    return q


def apply_dense(state, n, qubits, U):
    """U acts on `qubits` (qubits[b] = bit b of U's index), numpy reference."""
    m = len(qubits)
    psi = state.reshape([2] * n)                       # axis a <-> bit n-1-a
    axes = [n - 1 - q for q in reversed(qubits)]       # most significant matrix bit first
    psi = np.moveaxis(psi, axes, range(m))
    shp = psi.shape
    psi = (U @ psi.reshape(1 << m, -1)).reshape(shp)
    psi = np.moveaxis(psi, range(m), axes)
    return np.ascontiguousarray(psi).reshape(-1)


def pack_u(mats):
    out = []
    for U in mats:
        cm = np.asarray(U).T.reshape(-1)               # column-major
        out.append(np.stack([cm.real, cm.imag], axis=1).reshape(-1))
    return np.ascontiguousarray(np.concatenate(out))


CASES = [
    (12, [[0, 1, 2]]), (12, [[3, 7, 11]]), (13, [[5, 9, 10, 12]]), (14, [[0, 4, 8, 12, 13]]), (14, [[2, 3, 5, 7, 11, 13]]),
    (13, [[6]]), (13, [[1, 9]]), (15, [[9, 10, 11, 12, 13, 14]]), (14, [[3, 6, 9, 12]]),
    (16, [[4, 5, 6], [6, 7, 8, 9], [0, 15]]), (14, [[10, 11, 12, 13], [3, 10, 11, 12, 13]]), (12, [[4, 5, 6, 7, 8, 9], [0, 1, 2, 3, 10, 11]]),
]


@pytest.mark.parametrize("n,groups", CASES)
def test_dense_plan_tables_match_numpy(n, groups):
    rng = np.random.default_rng(n * 100 + len(groups))
    st = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    st /= np.linalg.norm(st)
    mats = [random_unitary(1 << len(q), rng) for q in groups]
    want = st.copy()
    for q, U in zip(groups, mats):
        want = apply_dense(want, n, q, U)
    m_list = (ctypes.c_int * len(groups))(*[len(q) for q in groups])
    flat = [b for q in groups for b in q]
    qpos = (ctypes.c_int * len(flat))(*flat)
    u = pack_u(mats)
    plan = ctypes.c_void_p()
    check(lib.hq_dense_plan_create(n, len(groups), m_list, qpos, u.ctypes.data, ctypes.byref(plan)))
    got = st.copy()
    check(lib.hq_debug_dense_plan_emulate(plan, got.ctypes.data))
    lib.hq_dense_plan_destroy(plan)
    assert np.max(np.abs(got - want)) < 1e-13


def test_dense_plan_rejects_bad_arguments():
    plan = ctypes.c_void_p()
    u = pack_u([np.eye(8)])
    m = (ctypes.c_int * 1)(3)
    assert lib.hq_dense_plan_create(12, 1, m, (ctypes.c_int * 3)(0, 1, 12), u.ctypes.data, ctypes.byref(plan)) != 0   # bit outside
    assert lib.hq_dense_plan_create(12, 1, m, (ctypes.c_int * 3)(4, 4, 5), u.ctypes.data, ctypes.byref(plan)) != 0    # repeated
    m7 = (ctypes.c_int * 1)(7)
    assert lib.hq_dense_plan_create(14, 1, m7, (ctypes.c_int * 7)(*range(7)), u.ctypes.data, ctypes.byref(plan)) != 0
