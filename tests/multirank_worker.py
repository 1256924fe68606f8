"""Worker of the world_size>1 CPU tests (gloo): every rank owns one shard of the state, replays its compiled schedule
with the plan emulator (test hook) and moves chunks between ranks exactly as the stage's SwapPlan says.  The GPU product
path does the same with hq_swap_begin / NCCL; this proves the partitioner's stage split, the per-rank / per-chunk gate
lowering, the swap semantics and the final layout bookkeeping without a GPU."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def swap_bits(idx, a, b):
    d = ((idx >> a) ^ (idx >> b)) & 1
    return idx ^ (d << a) ^ (d << b)


def run_rank(rank, world, port, text, out_path, env):
    os.environ.update(env)
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world)})
    import torch
    import torch.distributed as dist
    from hyquas_b200 import api
    from hyquas_b200._lib import check, lib

    dist.init_process_group("gloo", rank=rank, world_size=world)
    api.init_host_only(world, rank)
    c = api.Circuit.from_qasm(text)
    c.compile()
    n = c.num_qubits
    g = world.bit_length() - 1
    L = n - g
    shard = np.zeros(1 << L, dtype=np.complex128)
    if rank == 0:
        shard[0] = 1.0
    stages = lib.hq_debug_num_stages(c._h)
    I = ctypes.c_int
    total_overlap = 0
    for s in range(stages):
        npairs, k, nov = I(), I(), I()
        pa, pb, lb, gb = (I * 8)(), (I * 8)(), (I * 8)(), (I * 8)()
        check(lib.hq_debug_stage_swap(c._h, s, npairs, pa, pb, k, lb, gb, nov))
        total_overlap += nov.value
        kk = k.value
        if s > 0 and kk > 0:
            idx = np.arange(1 << L, dtype=np.int64)
            for i in range(npairs.value):
                idx = swap_bits(idx, pa[i], pb[i])
            shard = shard[idx]                       # product of disjoint transpositions = involution
            lbits = [lb[i] for i in range(kk)]
            if os.environ.get("HQ_TEST_SWAP_ANY", "0") != "1":
                assert lbits == list(range(L - kk, L))          # nccl transport: contiguous chunks
            allidx = np.arange(1 << L, dtype=np.int64)
            chunk_of = np.zeros(1 << L, dtype=np.int64)
            for i, b in enumerate(lbits):
                chunk_of |= ((allidx >> b) & 1) << i
            myc = sum(((rank >> gb[i]) & 1) << i for i in range(kk))
            for xr in range(1, 1 << kk):
                ch = myc ^ xr
                peer = rank
                for i in range(kk):
                    peer = (peer & ~(1 << gb[i])) | (((ch >> i) & 1) << gb[i])
                sel = chunk_of == ch                        # both sides enumerate the pair in order of the other L-k bits
                send = torch.from_numpy(np.ascontiguousarray(shard[sel]).view(np.float64).copy())
                recv = torch.empty_like(send)
                ops = [dist.P2POp(dist.isend, send, peer), dist.P2POp(dist.irecv, recv, peer)]
                for r in dist.batch_isend_irecv(ops):
                    r.wait()
                shard[sel] = recv.numpy().view(np.complex128)
            shard = np.ascontiguousarray(shard)
            for xr in range(1 << kk):                # arrival order of the product path: own chunk first
                check(lib.hq_debug_stage_emulate(c._h, s, 0, myc ^ xr, shard.ctypes.data))
        check(lib.hq_debug_stage_emulate(c._h, s, 1, 0, shard.ctypes.data))
    pos = (I * n)()
    check(lib.hq_debug_final_pos(c._h, pos))
    gathered = [torch.empty(2 << L, dtype=torch.float64) for _ in range(world)] if rank == 0 else None
    dist.gather(torch.from_numpy(shard.view(np.float64).copy()), gathered, dst=0)
    if rank == 0:
        phys = np.concatenate([t.numpy().view(np.complex128) for t in gathered])
        logical = np.arange(1 << n, dtype=np.int64)
        pid = np.zeros_like(logical)
        for q in range(n):
            pid |= ((logical >> q) & 1) << pos[q]
        np.save(out_path, phys[pid])
        with open(out_path + ".info", "w") as f:
            f.write(f"{stages} {total_overlap}")
    dist.barrier()
    dist.destroy_process_group()
    c.close()
