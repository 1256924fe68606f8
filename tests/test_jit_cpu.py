"""The per-group kernel specialiser (device/group_jit.cpp) on CPU: the emitter's HOST flavour -- same arithmetic text as the
CUDA source, threads and tiles replayed serially -- is compiled with g++ and compared with the oracle; the CUDA flavour is
compiled by NVRTC for sm_100a (no GPU needed) to prove it is valid source that fits the register budget."""
import ctypes
import os
import random
import re
import subprocess
import tempfile

import numpy as np
import pytest

from hyquas_b200 import circuits as C
from hyquas_b200._lib import check, lib
from oracle import oracle as O
from hyquas_b200._lib import HqGate


def pack(gates):
    arr = (HqGate * max(1, len(gates)))()
    for i, g in enumerate(gates):
        arr[i].type, arr[i].target, arr[i].control, arr[i].control2 = 0, g.target, g.control, g.control2
        m = np.asarray(g.mat).reshape(4)
        for j in range(4):
            arr[i].mat[2 * j], arr[i].mat[2 * j + 1] = m[j].real, m[j].imag
    return arr


def random_state(n, seed):
    r = np.random.default_rng(seed)
    s = (r.standard_normal(1 << n) + 1j * r.standard_normal(1 << n)).astype(np.complex128)
    return s / np.linalg.norm(s)


def jit_source(plan, host, zero_input=False):
    flavour = int(host) | (2 if zero_input else 0)
    need = ctypes.c_size_t()
    check(lib.hq_debug_group_plan_jit_source(plan, flavour, None, 0, ctypes.byref(need)))
    buf = ctypes.create_string_buffer(need.value)
    check(lib.hq_debug_group_plan_jit_source(plan, flavour, buf, need.value, ctypes.byref(need)))
    return buf.value.decode()


def run_host_flavour(src, state, amp0=0):
    with tempfile.TemporaryDirectory() as d:
        cpp, so = os.path.join(d, "k.cpp"), os.path.join(d, "k.so")
        open(cpp, "w").write(src)
        # -ffp-contract=off: the host build must not fuse what the source does not fuse
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-shared", "-fPIC", cpp, "-o", so])
        k = ctypes.CDLL(so)
        k.hq_group_jit_host.argtypes = [ctypes.c_void_p, ctypes.c_int]
        k.hq_group_jit_host(state.ctypes.data, amp0)


def make_plan(n, mask, gates):
    plan = ctypes.c_void_p()
    check(lib.hq_group_plan_create(n, mask, pack(gates), len(gates), ctypes.byref(plan)))
    return plan


@pytest.mark.parametrize("n,K,seed", [(12, 10, 0), (13, 11, 1), (14, 12, 2), (15, 12, 3), (12, 12, 4), (16, 11, 5)])
def test_specialised_source_matches_oracle(n, K, seed):
    rng = random.Random(seed)
    rest = list(range(3, n))
    rng.shuffle(rest)
    tile = [0, 1, 2] + sorted(rest[:K - 3])
    mask = sum(1 << b for b in tile)
    _, gates = O.parse_qasm(C.random_circuit(n, 250, seed))
    keep = [g for g in gates if (g.mat[0, 1] == 0 and g.mat[1, 0] == 0) or g.target in tile]
    plan = make_plan(n, mask, keep)
    st = random_state(n, seed)
    want = st.copy()
    O.apply(want, n, keep)
    run_host_flavour(jit_source(plan, True), st)
    lib.hq_group_plan_destroy(plan)
    assert np.max(np.abs(st - want)) < 1e-13


def test_specialised_source_supremacy_like_and_fixed_bits():
    """CZ / T / butterfly mixes (the sign-flip and deferred-scale paths) on a chunked launch (fixed bits)."""
    n, fixed_bit = 15, 13
    text = C.supremacy(n, cycles=12, seed=5)
    _, gates = O.parse_qasm(text)
    mask = 0xFFF
    keep = [g for g in gates if fixed_bit not in (g.target, g.control, g.control2)
            and ((g.mat[0, 1] == 0 and g.mat[1, 0] == 0) or g.target < 12)]
    st = random_state(n, 1)
    want = st.copy()
    O.apply(want, n, keep)
    for value in (0, 1):
        plan = ctypes.c_void_p()
        check(lib.hq_group_plan_create_ex(n, mask, 1 << fixed_bit, value << fixed_bit, pack(keep), len(keep), ctypes.byref(plan)))
        run_host_flavour(jit_source(plan, True), st)
        lib.hq_group_plan_destroy(plan)
    assert np.max(np.abs(st - want)) < 1e-13


def test_free_gates_emit_no_fp64():
    """X / Y / Z / S / CNOT / CZ on register qubits are renamings: the emitter's own instruction count stays 0."""
    gates = [O.OGate("x", 5), O.OGate("y", 6), O.OGate("z", 7), O.OGate("s", 8), O.OGate("cx", 6, 5), O.OGate("cz", 7, 8)]
    plan = make_plan(13, 0xFFF, gates)
    src = jit_source(plan, False)
    lib.hq_group_plan_destroy(plan)
    m = re.search(r"fp64 instructions per thread per tile: (\d+)", src)
    assert m and int(m.group(1)) == 0, src[-200:]


def test_cuda_flavour_compiles_for_sm100a(tmp_path):
    n = 14
    _, gates = O.parse_qasm(C.random_circuit(n, 120, 3))
    keep = [g for g in gates if (g.mat[0, 1] == 0 and g.mat[1, 0] == 0) or g.target < 12]
    plan = make_plan(n, 0xFFF, keep)
    src = jit_source(plan, False)
    lib.hq_group_plan_destroy(plan)
    log = ctypes.create_string_buffer(1 << 16)
    out = str(tmp_path / "k.cubin")
    rc = lib.hq_debug_jit_compile_to_file(src.encode(), out.encode(), log, len(log))
    if rc != 0 and b"libnvrtc not found" in log.value:
        pytest.skip("NVRTC not installed on this machine")
    assert rc == 0, log.value.decode()[:2000]
    info = log.value.decode()
    m = re.search(r"Used (\d+) registers", info)
    assert m and int(m.group(1)) <= 128, info
    sp = re.search(r"(\d+) bytes spill stores", info)   # 32 amplitudes' scalars + temporaries at the 128-register cap:
    assert sp is None or int(sp.group(1)) <= 512, info   # a few spilled temporaries are tolerated, a spilled working set is not
    assert os.path.getsize(out) > 1000


def test_block_fusion_shortens_su4_blocks_and_keeps_amplitudes():
    """Quantum-volume style blocks (u3 u3 / cx / u3 u3 / cx ...) on register qubits are emitted as 4x4 product matrices when that
    is cheaper: fewer FP64 instructions than gate by gate (HQ_JIT_NO_FUSE=1 in a child process), same amplitudes."""
    import subprocess
    import sys
    n = 13
    text = C.quantum_volume(n, depth=3, seed=11)
    _, gates = O.parse_qasm(text)
    keep = [g for g in gates if max(g.target, g.control) < 12]
    plan = make_plan(n, 0xFFF, keep)
    src = jit_source(plan, True)
    fused = int(re.search(r"fp64 instructions per thread per tile: (\d+)", src).group(1))
    assert int(re.search(r"(\d+) fused blocks", src).group(1)) > 0
    st = random_state(n, 5)
    want = st.copy()
    O.apply(want, n, keep)
    run_host_flavour(src, st)
    lib.hq_group_plan_destroy(plan)
    assert np.max(np.abs(st - want)) < 1e-13
    code = ("import sys, re; sys.path.insert(0, %r)\n"
            "from tests.test_jit_cpu import *\n"
            "_, gates = O.parse_qasm(C.quantum_volume(13, depth=3, seed=11))\n"
            "keep = [g for g in gates if max(g.target, g.control) < 12]\n"
            "src = jit_source(make_plan(13, 0xFFF, keep), False)\n"
            "print(re.search(r'fp64 instructions per thread per tile: (\\d+)', src).group(1))\n") % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, HQ_JIT_NO_FUSE="1"), timeout=120)
    assert r.returncode == 0, r.stderr[-500:]
    assert fused < 0.8 * int(r.stdout.strip())


@pytest.mark.parametrize("n,K,amp0", [(13, 10, 1), (14, 12, 1), (15, 12, 0)])
def test_zero_input_variant_needs_no_initialised_state(n, K, amp0):
    """The first group of a circuit acts on |0...0>: its zero-input variant reads nothing.  Run on a state full of garbage it must
    produce the group applied to |0...0> (amp0 = 1: this rank holds amplitude 0) or to the zero vector (amp0 = 0: another rank does)."""
    rng = random.Random(n)
    rest = list(range(3, n))
    rng.shuffle(rest)
    tile = [0, 1, 2] + sorted(rest[:K - 3])
    mask = sum(1 << b for b in tile)
    _, gates = O.parse_qasm(C.random_circuit(n, 120, n))
    keep = [g for g in gates if (g.mat[0, 1] == 0 and g.mat[1, 0] == 0) or g.target in tile]
    plan = make_plan(n, mask, keep)
    st = random_state(n, 77) * 1e3                      # garbage the kernel must never look at
    run_host_flavour(jit_source(plan, True, zero_input=True), st, amp0)
    lib.hq_group_plan_destroy(plan)
    want = np.zeros(1 << n, dtype=np.complex128)
    want[0] = float(amp0)
    O.apply(want, n, keep)
    assert np.max(np.abs(st - want)) < 1e-13


def test_zero_input_cuda_flavour_compiles(tmp_path):
    _, gates = O.parse_qasm(C.supremacy(14, cycles=8, seed=2))
    keep = [g for g in gates if (g.mat[0, 1] == 0 and g.mat[1, 0] == 0) or g.target < 12]
    plan = make_plan(14, 0xFFF, keep)
    src = jit_source(plan, False, zero_input=True)
    lib.hq_group_plan_destroy(plan)
    assert "HQ_ZERO_INPUT" in src and "cp.async.bulk" in src      # (the prologue text is shared; the variant's kernel body issues no loads)
    log = ctypes.create_string_buffer(1 << 16)
    rc = lib.hq_debug_jit_compile_to_file(src.encode(), str(tmp_path / "z.cubin").encode(), log, len(log))
    if rc != 0 and b"libnvrtc not found" in log.value:
        pytest.skip("NVRTC not installed on this machine")
    assert rc == 0, log.value.decode()[:2000]
    sass = subprocess.run(["cuobjdump", "-sass", str(tmp_path / "z.cubin")], capture_output=True, text=True).stdout
    assert "UBLKCP" not in sass and "STG" in sass


_CACHE_PROBE = r"""
import ctypes, sys
sys.path.insert(0, %(root)r)
from hyquas_b200._lib import lib
src = open(%(src)r).read().encode()
comp, hits = ctypes.c_int(), ctypes.c_int()
rc = lib.hq_debug_jit_cache_probe(%(ident)r, src, ctypes.byref(comp), ctypes.byref(hits))
print(rc, comp.value, hits.value)
"""


def test_kernel_cache_compiles_once_then_reads_the_disk(tmp_path):
    """The cache layer on its own (no GPU): a miss compiles and leaves <key>.cubin in $HQ_JIT_CACHE, a second process finds it there,
    another identity misses, and a damaged file is not accepted as a hit."""
    import subprocess
    import sys
    plan = make_plan(12, 0x3FF, [O.OGate("h", 8), O.OGate("cz", 8, 9), O.OGate("t", 9)])
    src = jit_source(plan, False)
    lib.hq_group_plan_destroy(plan)
    (tmp_path / "k.cu").write_text(src)
    cache = tmp_path / "cache" / "nested"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

    def probe(ident):
        code = _CACHE_PROBE % {"root": root, "src": str(tmp_path / "k.cu"), "ident": ident}
        out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, HQ_JIT_CACHE=str(cache)), capture_output=True,
                             text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        return [int(t) for t in out.stdout.split()[-3:]]

    first = probe(b"plan-a")
    if first[0] != 0:
        pytest.skip("NVRTC not installed on this machine")
    assert first == [0, 1, 0]
    files = sorted(cache.glob("*.cubin"))
    assert len(files) == 1 and files[0].read_bytes()[:4] == b"\x7fELF" and not list(cache.glob("*.tmp*"))
    assert probe(b"plan-a") == [0, 0, 1]            # second process: no compile, one disk hit
    assert probe(b"plan-b") == [0, 1, 0]            # the key follows the identity, not the source text
    assert len(list(cache.glob("*.cubin"))) == 2
    files[0].write_bytes(b"not a cubin")             # damaged entry: compiled again and replaced
    assert probe(b"plan-a") == [0, 1, 0]
    assert files[0].read_bytes()[:4] == b"\x7fELF"


_DUMP_KERNELS = r"""
import sys
sys.path.insert(0, %(root)r)
from hyquas_b200 import api, circuits as C
api.init_host_only(%(world)d, 0)
c = api.Circuit.from_qasm(C.generate(%(name)r))
c.compile()
print(len(c.groups()))
"""


@pytest.mark.parametrize("name,world", [("supremacy_30", 1), ("supremacy_31", 2)])
def test_every_kernel_of_the_headline_schedule_compiles_within_the_register_budget(tmp_path, name, world):
    """The schedule bench.py times (rank 0's share of it): every gate group's specialised kernel is emitted (HQ_JIT_DUMP_DIR, no GPU
    needed) and compiled for sm_100a here; none may fall back to the interpreter kernel on the GPU box because NVRTC rejects it,
    and none may spill its working set (512 threads per CTA: 128 registers)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", _DUMP_KERNELS % {"root": root, "name": name, "world": world}],
                         env=dict(os.environ, HQ_JIT_DUMP_DIR=str(tmp_path)), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    sources = sorted(tmp_path.glob("group_*.cu"))
    assert len(sources) >= int(out.stdout.split()[-1])
    log = ctypes.create_string_buffer(1 << 16)
    for src in sources:
        rc = lib.hq_debug_jit_compile_to_file(src.read_bytes(), str(src.with_suffix(".cubin")).encode(), log, len(log))
        if rc != 0 and b"libnvrtc not found" in log.value:
            pytest.skip("NVRTC not installed on this machine")
        info = log.value.decode()
        assert rc == 0, (src.name, info[:2000])
        m = re.search(r"Used (\d+) registers", info)
        assert m and int(m.group(1)) <= 128, (src.name, info)
        sp = re.search(r"(\d+) bytes spill stores", info)
        assert sp is None or int(sp.group(1)) <= 512, (src.name, info)
