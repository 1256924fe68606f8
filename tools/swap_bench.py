#!/usr/bin/env python
"""Bandwidth of the global<->local qubit swap alone (no gates): one process per GPU under torch.distributed.run.

    python -m torch.distributed.run --nproc-per-node 2 tools/swap_bench.py [L]

Prints GB/s per direction per GPU = 16 * 2^L * (1 - 2^-k) / time for the transports / CTA counts tried."""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from hyquas_b200 import api
    from hyquas_b200._lib import check, lib

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("gloo")
    api.init()
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    k = world.bit_length() - 1
    st = ctypes.c_void_p()
    check(lib.hq_state_alloc(L, ctypes.byref(st)))
    check(lib.hq_state_init(st, L, 1))
    check(lib.hq_swap_attach(st))
    anyp = ctypes.c_int()
    check(lib.hq_swap_any_position(anyp))
    I = ctypes.c_int
    out = {"L": L, "world": world, "k": k, "transport": "p2p" if anyp.value else "nccl", "runs": []}
    layouts = [("top", list(range(L - k, L)))]
    if anyp.value:
        layouts.append(("mid", list(range(12, 12 + k))))
        layouts.append(("low", list(range(5, 5 + k))))
    for label, lbits in layouts:
        for ctas in ([16, 32, 64, 148] if anyp.value else [0]):
            if ctas:
                os.environ["HQ_SWAP_CTAS"] = str(ctas)
            plan = ctypes.c_void_p()
            check(lib.hq_swap_plan_create(L, k, (I * k)(*lbits), (I * k)(*range(k)), ctypes.byref(plan)))
            best = 1e9
            for _ in range(3):
                check(lib.hq_sync())
                dist.barrier()
                t0 = time.perf_counter()
                check(lib.hq_swap_begin(plan, st))
                check(lib.hq_swap_end(plan))
                check(lib.hq_sync())
                best = min(best, time.perf_counter() - t0)
            check(lib.hq_swap_plan_destroy(plan))
            gbs = 16.0 * (1 << L) * (1 - 2.0 ** -k) / best / 1e9
            out["runs"].append({"bits": label, "ctas": ctas, "ms": round(best * 1e3, 3), "gbs_per_direction": round(gbs, 1)})
    check(lib.hq_swap_detach())
    check(lib.hq_state_free(st))
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
