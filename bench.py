#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json: circuit time & effective HBM TB/s, FP64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--circuit supremacy] [--qubits n]

One "step" = one execution of the workload circuit (all gate groups, i.e. S sweeps over the state).
Workload: configs[1] of BASELINE.json, `supremacy_30` on one B200 (16 GiB FP64 state); with N GPUs the state grows
to 30 + log2(N) qubits (weak scaling: 2^30 amplitudes per GPU).

  value   effective HBM TB/s = ALGORITHMIC bytes / device time, bytes = S * 32 * 2^L * N  (S = gate groups of the
          product's schedule for this circuit; 32 B = 16 read + 16 written per amplitude per sweep).
          Timed with CUDA events on the launching stream over exactly K back-to-back executions on the resident state.
  e2e     same metric through the public API from HOST inputs: QASM text -> parse -> compile (plans uploaded H2D) ->
          run (allocate, |0..0>, execute) -> amplitude dump read back D2H, wall clock per step.
  roofline      the kernel with the largest share of the step (tile kernel `group_kernel` or fused dense kernel
                `dense_kernel`): 32*2^L bytes / mean launch duration (CUDA events per launch) vs MEASURED_PEAKS.json
                hbm_gbs; the live-measured FP64 FMA / DMMA rates are reported next to it (gate-heavy launches are FP64-bound).
  cpu_baseline  the oracle's OpenMP gate-by-gate replay (kind "port") on a bounded sample of the same circuit.

`--impl reference` times the reference's own program (oracle/_ref/hyquas_ref_b*, built from /root/reference by
oracle/Makefile) on the same QASM file on this box's GPU, and falls back to the oracle replay on the host cores when
that binary is absent or cannot run the size.  Same metric, same byte numerator, so value ratios are time ratios.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ---------------------------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def workload(args, world):
    n = args.qubits or (30 + int(math.log2(world)))
    from hyquas_b200 import circuits
    name = f"{args.circuit}_{n}"
    return name, n, circuits.generate(name)


# ---------------------------------------------------------------------------------------------------------------
def cpu_replay(text: str, n: int, budget_s: float = 20.0):
    """Oracle replay of a bounded, evenly spaced sample of the circuit's gates at the full size n; returns
    (estimated seconds for the whole circuit, cores, description)."""
    from oracle import oracle as O
    _, gates = O.parse_qasm(text)
    cores = O.lib().orc_num_threads()
    # Bound the replay's footprint BEFORE touching memory (numpy's zero pages are lazy, so a MemoryError would come too late):
    # at most 30 qubits (16 GiB) and at most a quarter of what the host has free; the caller scales by 2^(n - n_used).
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 32 << 30
    n = min(n, 30)
    while n > 20 and (16 << n) > avail // 4:
        n -= 1
    state = O.zero_state(n)
    # spread the state so every gate does real work, then time a sample
    O.apply(state, n, [O.OGate("h", q) for q in range(min(n, 3))])
    per_gate_guess = (1 << n) * 16 * 2 / 8e9
    k = max(3, min(len(gates), int(budget_s / max(per_gate_guess, 1e-3))))
    step = max(1, len(gates) // k)
    sample = gates[::step][:k]
    t = O.apply(state, n, sample)
    est = t * len(gates) / len(sample)
    return est, cores, n, f"{len(sample)} of {len(gates)} gates (every {step}th) replayed at n={n}, scaled by gate count"


def run_reference(args, world, rank):
    """Reference arm: the reference's own binary on the same circuit (GPU), else the oracle replay (host cores)."""
    if rank != 0:
        return
    if world > 1:   # torchrun pins OMP_NUM_THREADS=1 per rank; only rank 0 works here, so it may use every host core
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    name, n, text = workload(args, world)
    peak, _ = measured_peaks()
    sweeps = int(os.environ.get("HQ_BENCH_SWEEPS", "0"))
    if not sweeps:   # numerator must equal the product arm's: take S from the product's partitioner (host only)
        from hyquas_b200 import api
        api.init_host_only(world, 0)
        c = api.Circuit.from_qasm(text)
        sweeps = c.plan_only()["groups"]
        c.close()
    bytes_per_step = sweeps * 32.0 * (1 << n)
    out = {"impl": "reference", "metric": "effective_hbm_tbps", "unit": "TB/s", "n_gpus": world,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": name, "qubits": n, "sweeps_counted": sweeps, "bytes_per_step": bytes_per_step}}
    times, kind, detail = [], None, None
    best_bin = None
    if world == 1 and n <= 30:
        qasm = os.path.join(tempfile.gettempdir(), f"{name}.qasm")
        open(qasm, "w").write(text)
        for backend, label in (("1", "group"), ("3", "blas")):
            exe = os.path.join(ROOT, "oracle", "_ref", f"hyquas_ref_b{backend}")
            if not os.path.exists(exe):
                continue
            try:
                r = subprocess.run([exe, qasm], capture_output=True, text=True, timeout=600)
                m = re.search(r"Time Cost: (\d+) us", r.stdout)
                if r.returncode == 0 and m:
                    t = int(m.group(1)) * 1e-6
                    if best_bin is None or t < best_bin[1]:
                        best_bin = (exe, t, label)
            except (subprocess.TimeoutExpired, OSError):
                pass
        if best_bin:
            exe, _, label = best_bin
            steps = max(1, min(args.steps, 3))
            for _ in range(steps):
                r = subprocess.run([exe, qasm], capture_output=True, text=True, timeout=600)
                times.append(int(re.search(r"Time Cost: (\d+) us", r.stdout).group(1)) * 1e-6)
            kind = "reference"
            detail = (f"reference's own CUDA build (backend {label}, sm_100, oracle/_ref) on this box's GPU; "
                      f"'Time Cost' of {len(times)} full runs of {name}; the reference has no CPU implementation")
            cores = 0
    if not times:
        est, cores, n_used, detail = cpu_replay(text, n)
        est *= 2.0 ** (n - n_used)
        times = [est]
        kind = "port"
        detail = "oracle OpenMP replay on host cores: " + detail
    sec = sum(times) / len(times)
    value = bytes_per_step / sec / 1e12
    out.update({"value": value, "steps": len(times), "warmup": 1 if best_bin else 0, "ms_per_step": sec * 1e3,
                "cpu_baseline": {"value": value, "unit": "TB/s", "cores": cores, "kind": kind, "sample": detail},
                "e2e": {"value": value, "unit": "TB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0})
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------------------------
def run_ours(args, world, rank, local_rank):
    import torch
    from hyquas_b200 import api

    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    api.init()
    name, n, text = workload(args, world)
    L = n - int(math.log2(world))
    peak, peak_src = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    c = api.Circuit.from_qasm(text)
    c.compile()
    info = c.schedule_info()
    S = info["groups"]
    bytes_per_step = S * 32.0 * (1 << L) * world
    c.prepare_state()

    # ---- value: K executions on the resident state, CUDA events on the launching stream -------------------
    for _ in range(max(3, args.warmup)):
        c.execute()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_ms = 0.0
    for _ in range(args.steps):
        _, ms = c.execute()
        dev_ms += ms
    barrier()
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([dev_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = bytes_per_step / (ms_per_step * 1e-3) / 1e12

    # ---- roofline: per-launch durations (CUDA events on the compute stream) of the kernel with the largest time share ----
    _, _, per_group = c.execute(per_group=True)
    per_group2 = c.execute(per_group=True)[2]
    per_launch = [min(a, b) for a, b in zip(per_group, per_group2)] if len(per_group) == len(per_group2) else per_group
    ginfo = c.groups()
    # execute() reports one entry per LAUNCH; a per-chunk group has several: fold them back per group
    groups, pos = [], 0
    for g in ginfo:
        ms = sum(per_launch[pos:pos + g["launches"]])
        pos += g["launches"]
        groups.append({"backend": g["backend"], "gates": g["gates"], "blocks": g["blocks"], "ms": round(ms, 3),
                       "predicted_ms": round(g["predicted_ms"], 3)})
    share = {}
    for g in groups:
        share[g["backend"]] = share.get(g["backend"], 0.0) + g["ms"]
    dominant = max(share, key=share.get) if share else "tile"
    dom = [g["ms"] for g in groups if g["backend"] == dominant]
    mean_launch_ms = sum(dom) / max(1, len(dom))
    alg_bytes = 32.0 * (1 << L)
    achieved = alg_bytes / (mean_launch_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(f"{dominant}_kernel_dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "group_kernel" if dominant == "tile" else "dense_kernel", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "launch_ms_mean": mean_launch_ms, "launch_ms_min": min(dom) if dom else None,
                "launch_ms_max": max(dom) if dom else None, "time_share": share.get(dominant, 0.0) / max(1e-9, sum(share.values())),
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "one launch = one in-place sweep of the local state; launches that carry many gates are FP64-bound "
                        "(see fp64 below), so frac < 1 is arithmetic, not wasted traffic"}
    import ctypes
    from hyquas_b200._lib import check, lib
    v = ctypes.c_double()
    check(lib.hq_microbench_fp64(0, v)); fma_tf = v.value
    check(lib.hq_microbench_fp64(1, v)); mma_tf = v.value
    roofline["fp64"] = {"fma_tflops_measured": fma_tf, "dmma_tflops_measured": mma_tf,
                        "source": "hq_microbench_fp64, run live in this process (MEASURED_PEAKS.json has no FP64 figure)"}
    c.close()

    # ---- e2e: the public API from host inputs (QASM text) to host outputs (amplitude dump) -----------------
    e2e_steps = max(1, min(args.steps, 3))
    h2d = d2h = 0
    parts = {"parse": 0.0, "compile_and_plan_upload": 0.0, "run_alloc_init_execute_dump_readback": 0.0, "free": 0.0}
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ta = time.perf_counter()
        ce = api.Circuit.from_qasm(text)
        tb = time.perf_counter()
        ce.compile()
        tc = time.perf_counter()
        ce.run(copy_back=False, destroy=False)
        dump = ce.dump()
        td = time.perf_counter()
        h2d, d2h = ce.io_bytes()
        ce.close()
        te = time.perf_counter()
        for k, v in zip(parts, (tb - ta, tc - tb, td - tc, te - td)):
            parts[k] += v * 1e3 / e2e_steps
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": bytes_per_step / e2e_s / 1e12, "unit": "TB/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
           "breakdown_ms": {k: round(v, 2) for k, v in parts.items()},
           "path": "QASM text -> hq_circuit_from_qasm -> compile -> run(alloc, init, execute) -> dump"}

    # ---- cpu baseline (rank 0, N=1 only) --------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        est, cores, n_used, detail = cpu_replay(text, n)
        est *= 2.0 ** (n - n_used)
        cpu = {"value": bytes_per_step / est / 1e12, "unit": "TB/s", "cores": cores, "kind": "port",
               "sample": detail, "est_circuit_seconds": est}

    if rank == 0:
        out = {"metric": "effective_hbm_tbps", "value": value, "unit": "TB/s", "n_gpus": world, "steps": args.steps,
               "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": name, "qubits": n, "local_qubits": L, "gates": info["gates"], "sweeps": S,
                          "stages": info["stages"], "bytes_per_step": bytes_per_step,
                          "state_bytes_per_gpu": 16 * (1 << L), "l2": "inputs (16 GiB state) larger than L2",
                          "tile_bits": int(os.environ.get("HQ_TILE_BITS", "12")),
                          "backend": os.environ.get("HQ_BACKEND", "mix")},
               "circuit_time_ms": ms_per_step, "sweeps_per_s": S / (ms_per_step * 1e-3),
               "roofline": roofline, "groups": groups, "cpu_baseline": cpu, "e2e": e2e,
               "gpu_launches": sum(g["launches"] for g in ginfo) * args.steps,
               "clocks": clocks}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--circuit", default="supremacy")
    ap.add_argument("--qubits", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, world, rank)
    else:
        run_ours(args, world, rank, local_rank)


if __name__ == "__main__":
    main()
