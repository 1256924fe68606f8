"""world_size-2/4/8 `gloo` tests of the multi-GPU path's host logic (no GPU): stage split, per-rank lowering, swap plan,
per-chunk overlap groups, final layout.  See tests/multirank_worker.py for what each rank does."""
import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch.multiprocessing as mp

from hyquas_b200 import circuits as C
from oracle import oracle as O
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from multirank_worker import run_rank  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, text, env=None):
    out = os.path.join(tempfile.mkdtemp(), "state.npy")
    mp.spawn(run_rank, args=(world, _free_port(), text, out, env or {}), nprocs=world, join=True)
    stages, overlap = map(int, open(out + ".info").read().split())
    return np.load(out), stages, overlap


@pytest.mark.parametrize("any_bit", ["0", "1"])
@pytest.mark.parametrize("world,name", [(2, "qft_14"), (2, "supremacy_14"), (2, "adder_14"),
                                        (4, "quantum_volume_14"), (4, "hidden_shift_14"), (4, "basis_change_14"),
                                        (4, "bv_15"), (8, "qaoa_16"), (8, "adder_16"), (8, "supremacy_15")])
def test_sharded_schedule_matches_oracle(world, name, any_bit):
    """any_bit=0: swaps trade the top k local positions (nccl transport); 1: any position >= 5 (p2p transport)."""
    text = C.generate(name)
    got, stages, _ = _run(world, text, env={"HQ_TEST_SWAP_ANY": any_bit})
    n, gates = O.parse_qasm(text)
    want = O.simulate(n, gates)
    assert np.max(np.abs(got - want)) <= 1e-10
    assert stages >= 1


def test_random_circuit_needs_several_exchanges():
    names = ["h", "x", "y", "z", "s", "sdg", "t", "tdg", "rx", "ry", "rz", "u1", "u3", "cx", "cy", "cz", "crx", "cry",
             "crz", "cu1", "ccx"]
    text = C.random_circuit(14, 400, seed=21, names=names)
    got, stages, overlap = _run(4, text, env={"HQ_TEST_SWAP_ANY": "1"})
    n, gates = O.parse_qasm(text)
    assert np.max(np.abs(got - O.simulate(n, gates))) <= 1e-10
    assert stages >= 3          # every qubit is a non-diagonal target many times: the layout must keep rotating


@pytest.mark.parametrize("any_bit", ["0", "1"])
@pytest.mark.parametrize("world,name", [(2, "supremacy_14"), (4, "qaoa_14"), (4, "quantum_volume_14"), (2, "adder_14")])
def test_overlap_groups_present_and_optional(world, name, any_bit):
    """Per-chunk (overlap) groups: forced on with a huge slack (at 14 qubits the predicted exchange is too short to hide
    anything), and switched off -- same amplitudes either way."""
    text = C.generate(name)
    n, gates = O.parse_qasm(text)
    want = O.simulate(n, gates)
    got, _, overlap = _run(world, text, env={"HQ_OVERLAP_SLACK": "1e9", "HQ_OVERLAP_MODE": "split", "HQ_TEST_SWAP_ANY": any_bit})
    assert overlap >= 1 and np.max(np.abs(got - want)) <= 1e-10
    got, _, overlap = _run(world, text, env={"HQ_ENABLE_OVERLAP": "0", "HQ_TEST_SWAP_ANY": any_bit})
    assert overlap == 0 and np.max(np.abs(got - want)) <= 1e-10
