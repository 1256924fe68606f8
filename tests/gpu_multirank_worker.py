"""Launched by tests/test_gpu_multi.py under torch.distributed.run (one process per GPU): runs circuits through the
product path (NCCL swap + per-chunk overlap groups + full groups) and checks rank-assembled amplitudes and the dump
text against the oracle on rank 0.  Exit code != 0 on any mismatch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from hyquas_b200 import api, circuits as C
    from oracle import oracle as O

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("gloo")
    api.init()
    names = sys.argv[1:] or ["qft_20", "supremacy_20", "qaoa_20", "adder_20", "quantum_volume_18", "hidden_shift_20",
                             "basis_change_18", "bv_20", "supremacy_24"]
    bad = 0
    for name in names:
        text = C.generate(name)
        c = api.Circuit.from_qasm(text)
        c.compile()
        c.run(copy_back=False, destroy=False)
        n = c.num_qubits
        shard = c.local_shard(world)
        pos = c.final_layout()
        dump = c.dump()
        info = c.schedule_info()
        # ampAt is collective (broadcast from the owner): every rank asks for the same indices and gets the same values
        probes = [0, 5, (1 << n) - 1, (1 << (n - 1)) + 77]
        amps = [c.amp_at(i) for i in probes]
        p0 = [c.measure(q) for q in (0, n // 2, n - 1)]           # collective too
        gathered = [torch.empty(shard.size * 2, dtype=torch.float64) for _ in range(world)] if rank == 0 else None
        dist.gather(torch.from_numpy(shard.view(np.float64).copy()), gathered, dst=0)
        if rank == 0:
            phys = np.concatenate([t.numpy().view(np.complex128) for t in gathered])
            logical = np.arange(1 << n, dtype=np.int64)
            pid = np.zeros_like(logical)
            for q in range(n):
                pid |= ((logical >> q) & 1) << pos[q]
            got = phys[pid]
            _, gates = O.parse_qasm(text)
            want = O.simulate(n, gates)
            err = float(np.max(np.abs(got - want)))
            ok, derr = O.compare_dumps(O.dump_state(want, n), dump)
            ok = ok and all(abs(a - want[i]) <= 1e-10 for a, i in zip(amps, probes))
            for q, p in zip((0, n // 2, n - 1), p0):
                ok = ok and abs(p - float(np.sum(np.abs(want[((logical >> q) & 1) == 0]) ** 2))) <= 1e-10
            status = "ok" if err <= 1e-10 and ok else "MISMATCH"
            bad += status != "ok"
            print(f"[multi-gpu x{world}] {name}: stages={info['stages']} groups={info['groups']} max|d|={err:.2e} dump={ok} {status}",
                  flush=True)
        c.close()
    flag = torch.tensor([bad])
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
