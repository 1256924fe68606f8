#!/usr/bin/env python
"""B200 microbenchmarks of the gate-group kernel: synthetic single-purpose circuits, per-launch CUDA-event times.

    python tools/microbench.py [--qubits 30] [--out gpurun_out/microbench.json]

Each case is a tiny circuit built through the public API; the numbers are ms per gate-group launch over 2^n amplitudes.
Used (a) to calibrate the evaluator (tools/calibrate.py reads the JSON) and (b) as kernel A/B evidence.
"""
import argparse
import ctypes
import json
import math
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hyquas_b200 import api  # noqa: E402
from hyquas_b200._lib import check, lib  # noqa: E402


def build(n, spec, count, qubits, seed=1):
    rng = random.Random(seed)
    c = api.Circuit(n)
    for i in range(count):
        name = spec[i % len(spec)]
        q = qubits[i % len(qubits)]
        if name in ("H", "X", "Y", "Z", "S", "T", "SDG", "TDG"):
            c.add_gate(name, q)
        elif name in ("RXH", "RYH"):       # the supremacy circuits' rx(pi/2), ry(pi/2): butterfly class
            c.add_gate(name[:2], q, params=(math.pi / 2,))
        elif name in ("RX", "RY", "RZ", "U1"):
            c.add_gate(name, q, params=(rng.uniform(0.1, 3.0),))
        elif name == "U3":
            c.add_gate(name, q, params=(rng.uniform(0.1, 3.0), rng.uniform(0.1, 3.0), rng.uniform(0.1, 3.0)))
        elif name in ("CZ", "CNOT", "CY"):
            q2 = qubits[(i + 1) % len(qubits)]
            c.add_gate(name, q, q2)
        elif name in ("CU1F", "CZF"):     # fan: one operand on the hot qubits, the partner elsewhere (QFT ladders, CZ fans)
            q2 = 12 + (i * 7) % (n - 12)
            if name == "CZF":
                c.add_gate("CZ", q, q2)
            else:
                c.add_gate("CU1", q, q2, params=(rng.uniform(0.1, 3.0),))
        elif name in ("CRX", "CRY", "CRZ", "CU1"):
            q2 = qubits[(i + 1) % len(qubits)]
            c.add_gate(name, q, q2, params=(rng.uniform(0.1, 3.0),))
        else:
            raise ValueError(name)
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--out", default="")
    ap.add_argument("--only", default="", help="comma-separated tile-kernel case names; skips the FP64/copy/dense sections")
    args = ap.parse_args()
    only = [x for x in args.only.split(",") if x]
    n = args.qubits
    api.init()
    res = {"qubits": n, "lib_suffix": os.environ.get("HQ_LIB_SUFFIX", ""), "cases": {}}
    v = ctypes.c_double()
    for kind, key in [] if only else ((0, "fp64_fma_tflops"), (1, "fp64_mma_tflops")):
        check(lib.hq_microbench_fp64(kind, v))
        res[key] = v.value
    res["dmma_tflops_by_warps_per_sm"] = {}
    for w in () if only else (4, 8, 12, 16, 24, 32):
        check(lib.hq_microbench_fp64(10 + w, v))
        res["dmma_tflops_by_warps_per_sm"][w] = v.value
    if not only:
        print("dmma by warps/SM:", {k: round(x, 1) for k, x in res["dmma_tflops_by_warps_per_sm"].items()}, flush=True)
    st = ctypes.c_void_p()
    check(lib.hq_state_alloc(n, ctypes.byref(st)))
    check(lib.hq_microbench_copy(st, n, v))
    res["copy_gbs"] = v.value
    check(lib.hq_state_free(st))
    print(json.dumps({k: res.get(k) for k in ("fp64_fma_tflops", "fp64_mma_tflops", "copy_gbs")}), flush=True)

    hi4 = [8, 9, 10, 11]
    cases = {
        "sweep_1gate": (["H"], 1, [8]),
        "h_x64_4q": (["H"], 64, hi4),
        "rx_x64_4q": (["RX"], 64, hi4),
        "u3_x64_4q": (["U3"], 64, hi4),
        "t_x64_4q": (["T"], 64, hi4),
        "rz_x64_4q": (["RZ"], 64, hi4),
        "x_x64_4q": (["X"], 64, hi4),
        "cz_x64_4q": (["CZ"], 64, hi4),
        "cnot_x64_4q": (["CNOT"], 64, hi4),
        "cu1fan_x64_4q": (["CU1F"], 64, hi4),
        "czfan_x64_4q": (["CZF"], 64, hi4),
        "h_cu1fan_x64_4q": (["H", "CU1F", "CU1F", "CU1F", "CU1F", "CU1F", "CU1F", "CU1F"], 64, hi4),
        "mix5_x64_4q": (["H", "RX", "T", "RY", "CZ"], 64, hi4),
        "h_x64_1q": (["H"], 64, [9]),
        "h_x96_12q": (["H"], 96, list(range(12))),
        "mix5_x96_12q": (["H", "RX", "T", "RY", "CZ"], 96, list(range(12))),
        "u3_x96_12q": (["U3"], 96, list(range(12))),
        "h_x16_4q": (["H"], 16, hi4),
        "h_x256_4q": (["H"], 256, hi4),
    }
    os.environ["HQ_BACKEND"] = "group"      # these cases price the TILE kernel; the hybrid partitioner would fuse them
    os.environ["HQ_PEEPHOLE_MERGE"] = "0"   # ... and the merge pass would multiply the repeated gates together
    cases.update({
        "h_x96_12q_hi": (["H"], 96, list(range(16, 28))),     # same, on qubits the partitioner is free to place
        "t_x96_12q": (["T"], 96, list(range(12))),
        "sup5_x64_4q": (["H", "RXH", "T", "RYH", "CZ"], 64, hi4),
        "sup5_x96_12q": (["H", "RXH", "T", "RYH", "CZ"], 96, list(range(12))),
        "mix5_x256_12q": (["H", "RX", "T", "RY", "CZ"], 256, list(range(12))),
    })
    # tile geometry: the same light / medium work on tiles made of the 5 pinned low bits + 7 qubits at different heights
    # (contiguous 64 KB tile ... 128 runs of 512 B spread over the whole state)
    for tag, qs in (("lo7", list(range(5, 12))), ("mid7", list(range(12, 19))), ("hi7", list(range(n - 8, n - 1))),
                    ("spread7", [6, 9, 13, 17, 21, 25, n - 1])):
        cases["sweep_" + tag] = (["H"], 7, qs)
        cases["sup5_x42_" + tag] = (["H", "RXH", "T", "RYH", "CZ"], 42, qs)
        cases["sup5_x84_" + tag] = (["H", "RXH", "T", "RYH", "CZ"], 84, qs)
    for name, (spec, count, qubits) in cases.items():
        if only and name not in only:
            continue
        c = build(n, spec, count, qubits)
        c.compile()
        c.prepare_state()
        for _ in range(2):
            c.execute()
        _, ms, per = c.execute(per_group=True)
        _, ms2, per2 = c.execute(per_group=True)
        per = [min(a, b) for a, b in zip(per, per2)]
        res["cases"][name] = {"gates": count, "groups": len(per), "ms_total": min(ms, ms2), "per_group_ms": per}
        print(f"{name:16s} gates={count:4d} groups={len(per)} total={min(ms, ms2):8.3f} ms  per-group={['%.2f' % x for x in per]}",
              flush=True)
        c.close()
    os.environ.pop("HQ_BACKEND", None)
    if only:
        if args.out:
            json.dump(res, open(args.out, "w"), indent=1)
        return
    # fused dense kernel: one or several random unitaries on m qubits, 2^n amplitudes
    import numpy as np
    rng = np.random.default_rng(5)
    st = ctypes.c_void_p()
    check(lib.hq_state_alloc(n, ctypes.byref(st)))
    check(lib.hq_state_init(st, n, 1))
    res["dense"] = {}
    dense_cases = {"m3_low": [[0, 1, 2]], "m3_high": [[20, 21, 22]], "m4_low": [[0, 1, 2, 3]], "m4_high": [[20, 21, 22, 23]],
                   "m5_low": [[0, 1, 2, 3, 4]], "m5_high": [[20, 21, 22, 23, 24]], "m6_low": [[0, 1, 2, 3, 4, 5]],
                   "m6_high": [[20, 21, 22, 23, 24, 25]], "m4_mixed": [[1, 9, 17, 25]],
                   "m4x2": [[3, 4, 5, 6], [7, 8, 9, 10]], "m4x3": [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11]], "m3x3_high": [[20, 21, 22], [23, 24, 25], [26, 27, 28]],
                   "m3x2": [[3, 4, 5], [6, 7, 8]], "m5x2": [[0, 1, 2, 3, 4], [5, 6, 7, 8, 9]], "m3x8": [[3 + (i % 3) * 3 + j for j in range(3)] for i in range(8)],
                   "m3x3": [[3, 4, 5], [6, 7, 8], [9, 10, 11]], "m2x4": [[3, 4], [5, 6], [7, 8], [9, 10]]}
    for name, groups in dense_cases.items():
        mats = []
        for q in groups:
            a = rng.standard_normal((1 << len(q), 1 << len(q))) + 1j * rng.standard_normal((1 << len(q), 1 << len(q)))
            U = np.linalg.qr(a)[0]
            cm = U.T.reshape(-1)
            mats.append(np.stack([cm.real, cm.imag], axis=1).reshape(-1))
        u = np.ascontiguousarray(np.concatenate(mats))
        flat = [b for q in groups for b in q]
        plan = ctypes.c_void_p()
        check(lib.hq_dense_plan_create(n, len(groups), (ctypes.c_int * len(groups))(*[len(q) for q in groups]),
                                       (ctypes.c_int * len(flat))(*flat), u.ctypes.data, ctypes.byref(plan)))
        best = 1e9
        for _ in range(4):
            check(lib.hq_timer_start())
            check(lib.hq_dense_plan_launch(plan, st, 0))
            ms = ctypes.c_float()
            check(lib.hq_timer_stop_ms(ms))
            best = min(best, ms.value)
        fl = ctypes.c_double()
        check(lib.hq_dense_plan_info(plan, None, None, None, fl, None))
        lib.hq_dense_plan_destroy(plan)
        res["dense"][name] = {"ms": best, "gbs": 32.0 * (1 << n) / best / 1e6, "tflops": fl.value * (1 << n) / best / 1e9}
        print(f"dense {name:10s} {best:8.3f} ms  {res['dense'][name]['gbs']:7.0f} GB/s  {res['dense'][name]['tflops']:6.2f} TFLOP/s", flush=True)
    check(lib.hq_state_free(st))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
