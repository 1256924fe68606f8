#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json: circuit time & effective HBM TB/s, FP64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--circuit supremacy] [--qubits n]

One "step" = one execution of the workload circuit (all gate groups, i.e. S sweeps over the state).
Workload: configs[1] of BASELINE.json, `supremacy_30` on one B200 (16 GiB FP64 state); with N GPUs the state grows
to 30 + log2(N) qubits (weak scaling: 2^30 amplitudes per GPU).

  value   effective HBM TB/s = NOMINAL bytes / device time.  The numerator is frozen per workload (nominal_sweeps():
          the sweep count of the round-1 schedule of the same circuit, else gates / 40), times 32 B per amplitude per
          sweep, so that neither arm's partitioner can move the score: for a given workload `value` is inversely
          proportional to circuit time, in both arms and across rounds.  `circuit_time_ms` is reported next to it and
          `value_executed_schedule` uses the sweeps the product actually ran (SURVEY.md 8(d)'s definition).
          Timed with CUDA events on the launching stream over exactly K back-to-back executions on the resident state.
  e2e     same metric through the public API from HOST inputs: QASM text -> parse -> compile (plans uploaded H2D, kernels
          fetched from the JIT cache) -> run (allocate, |0..0>, execute) -> amplitude dump read back D2H, wall clock per
          step.  The first e2e step of a process compiles this circuit's specialised kernels (NVRTC, all cores) unless the
          on-disk cache is warm; that one-off is reported as `jit` and is not part of the steady-state figure.
  roofline      the kernel with the largest share of the step: 32*2^L bytes / mean launch duration (CUDA events per
                launch) vs MEASURED_PEAKS.json hbm_gbs; live-measured FP64 rates are reported next to it.
  cpu_baseline  the oracle's OpenMP gate-by-gate replay (kind "port") on >= 32 gates of the same circuit, stratified by gate type.
  parity        before timing, small instances of the workload family run through the same product path and are compared
                with the oracle on rank 0 (max |delta amp| <= 1e-10, dump text equal); a mismatch exits non-zero.

`--impl reference` times the reference's own program (oracle/_ref/hyquas_ref_b*, built from /root/reference by
oracle/Makefile: `group`, `blas` and `mix` backends; `mix` reads parameter files that the reference's own preprocess tool
writes on this box) on the same QASM file, on GPUs 0..N-1 of this box (CUDA_VISIBLE_DEVICES pinned; N > 1 = the reference's
single-process peer-copy mode).  It imports nothing from hyquas_b200 except the circuit generator (pure Python).
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

# Sweep counts of the round-1 schedules (BENCH_r01.json / SCALE_r01.json), frozen as the numerator of `value`.
NOMINAL_SWEEPS = {"supremacy": {30: 13, 31: 17, 32: 16, 33: 16}}


def count_gates(text: str) -> int:
    return sum(1 for l in text.splitlines() if l and not l.startswith(("OPENQASM", "include", "qreg", "//")))


def nominal_sweeps(name: str, n: int, text: str) -> int:
    fam = name.rsplit("_", 1)[0]
    return NOMINAL_SWEEPS.get(fam, {}).get(n) or max(1, round(count_gates(text) / 40))


def load_circuits():
    """hyquas_b200/circuits.py by path: the generators are pure Python and must not pull the product library into the reference arm."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("hq_circuits", os.path.join(ROOT, "hyquas_b200", "circuits.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


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
    name = f"{args.circuit}_{n}"
    return name, n, load_circuits().generate(name)


# ---------------------------------------------------------------------------------------------------------------
def cpu_replay(text: str, n: int, budget_s: float = 20.0, min_gates: int = 32):
    """Oracle replay of a bounded sample of the circuit's gates at (up to) the full size n, STRATIFIED by gate name: every
    gate type that occurs is sampled in proportion to its count (at least once, evenly spaced), at least `min_gates` gates in
    all; the circuit estimate is sum over types of (mean sampled time of the type) x (count of the type).
    Returns (estimated seconds for the whole circuit at n_used, cores, n_used, description)."""
    from oracle import oracle as O
    _, gates = O.parse_qasm(text)
    cores = O.lib().orc_num_threads()
    # Bound the replay's footprint BEFORE touching memory: at most 30 qubits (16 GiB) and a quarter of what the host has free
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 32 << 30
    n = min(n, 30)
    while n > 20 and (16 << n) > avail // 4:
        n -= 1
    state = O.zero_state(n)
    O.apply(state, n, [O.OGate("h", q) for q in range(min(n, 3))])   # spread the state so every gate does real work
    t_probe = O.apply(state, n, gates[:2]) / 2                        # seconds per gate on this box
    k = int(max(min_gates, min(len(gates), budget_s / max(t_probe, 1e-4))))
    by_type = {}
    for g in gates:
        by_type.setdefault(g.name, []).append(g)
    est, used, parts = 0.0, 0, []
    for name, lst in sorted(by_type.items()):
        take = max(1, min(len(lst), round(k * len(lst) / len(gates))))
        step = len(lst) / take
        sample = [lst[int(i * step)] for i in range(take)]
        t = O.apply(state, n, sample)
        est += t / take * len(lst)
        used += take
        parts.append(f"{name}:{take}/{len(lst)}")
    return est, cores, n, (f"{used} of {len(gates)} gates replayed at n={n}, stratified by gate type ({' '.join(parts)}); "
                           f"estimate = sum over types of mean time x count")


# ---------------------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_BACKENDS = (("1", "group"), ("3", "blas"), ("4p", "mix"))


def ref_run_once(exe, qasm, env, timeout=900):
    """One run of a reference binary -> (Time Cost seconds or None, wall seconds, stdout tail)."""
    t0 = time.perf_counter()
    try:
        r = subprocess.run([exe, qasm], capture_output=True, text=True, timeout=timeout, env=env,
                           cwd=os.path.join(REF_DIR, "run"))
    except (subprocess.TimeoutExpired, OSError) as e:
        return None, time.perf_counter() - t0, str(e)
    m = re.search(r"Time Cost: (\d+) us", r.stdout)
    ok = r.returncode == 0 and m
    return (int(m.group(1)) * 1e-6 if ok else None), time.perf_counter() - t0, (r.stdout + r.stderr)[-400:]


def ref_parameter_files(local_qubits, env, log):
    """`mix` needs ../evaluator-preprocess/parameter-files/{L}qubits.out (src/evaluator.cpp:60-103).  The reference's own
    preprocess tool (oracle/_ref/hyquas_ref_process = evaluator-preprocess/process.cpp, qubit list from argv) writes it; on a
    B200 that takes 11 minutes for L = 30 (2000 cuTT plans + 1000 Zgemm calls), so the file it wrote on this pool's B200 in
    session r02_s2 is kept under oracle/ref_params/ and copied into place; HQ_REF_GENERATE_PARAMS=1 regenerates it on GPU 0."""
    pdir = os.path.join(REF_DIR, "evaluator-preprocess", "parameter-files")
    os.makedirs(pdir, exist_ok=True)
    os.makedirs(os.path.join(REF_DIR, "run"), exist_ok=True)
    path = os.path.join(pdir, f"{local_qubits}qubits.out")
    tool = os.path.join(REF_DIR, "hyquas_ref_process")
    if os.path.exists(path) and os.path.getsize(path) > 100:
        return True
    kept = os.path.join(ROOT, "oracle", "ref_params", f"{local_qubits}qubits.out")
    if os.path.exists(kept):
        import shutil
        shutil.copyfile(kept, path)
        log.append(f"parameter file for L={local_qubits}: oracle/ref_params (written by the reference's process tool on this pool's B200)")
        return True
    if not os.path.exists(tool) or os.environ.get("HQ_REF_GENERATE_PARAMS") != "1":
        return False
    e = dict(env, CUDA_VISIBLE_DEVICES="0")
    t0 = time.perf_counter()
    try:
        r = subprocess.run([tool, str(local_qubits)], capture_output=True, text=True, timeout=1500, env=e, cwd=os.path.join(REF_DIR, "run"))
        log.append(f"process {local_qubits}: rc={r.returncode} {time.perf_counter() - t0:.0f}s")
        return r.returncode == 0 and os.path.exists(path) and os.path.getsize(path) > 100
    except (subprocess.TimeoutExpired, OSError) as ex:
        log.append(f"process {local_qubits}: {ex}")
        return False


def run_reference(args, world, rank):
    """Reference arm: the reference's own binary on the same circuit and the same number of GPUs of this box."""
    if rank != 0:
        return
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)   # torchrun pins 1 per rank; only rank 0 works here
    name, n, text = workload(args, world)
    sweeps = nominal_sweeps(name, n, text)
    L = n - int(math.log2(world))
    bytes_per_step = sweeps * 32.0 * (1 << n)
    out = {"impl": "reference", "metric": "effective_hbm_tbps", "unit": "TB/s", "n_gpus": world,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": name, "qubits": n, "local_qubits": L, "nominal_sweeps": sweeps, "bytes_per_step": bytes_per_step,
                      "numerator": "frozen per workload (nominal_sweeps x 32 B x 2^n), identical in both arms"}}
    env = dict(os.environ, CUDA_VISIBLE_DEVICES=",".join(str(i) for i in range(world)))
    env.pop("HQ_LIB_SUFFIX", None)
    qasm = os.path.join(tempfile.gettempdir(), f"{name}.qasm")
    open(qasm, "w").write(text)
    notes, cands = [], []
    can_run = L <= 30 or (L <= 31 and world <= 8)   # 2 x 16 B x 2^L per GPU must fit (reference doubles the buffer), 32-bit masks
    if can_run and os.path.isdir(REF_DIR):
        os.makedirs(os.path.join(REF_DIR, "run"), exist_ok=True)
        have_params = ref_parameter_files(L, env, notes)
        for b, label in REF_BACKENDS:
            exe = os.path.join(REF_DIR, f"hyquas_ref_b{b}")
            if not os.path.exists(exe) or (b == "4p" and not have_params):
                notes.append(f"{label}: not available")
                continue
            t, wall, tail = ref_run_once(exe, qasm, env)
            if t is None:
                notes.append(f"{label}: failed ({tail[-120:].strip()})")
            else:
                cands.append((t, wall, exe, label))
                notes.append(f"{label}: {t * 1e3:.1f} ms")
    times, kind, cores = [], None, 0
    if cands:
        # the reference's best backend on this circuit is the baseline; every candidate's first run doubles as warm-up
        cands.sort()
        _, wall, exe, label = cands[0]
        budget = 240.0
        warm = max(0, min(args.warmup - 1, int(budget * 0.2 / max(wall, 0.1))))
        steps = max(1, min(args.steps, int(budget * 0.8 / max(wall, 0.1))))
        for _ in range(warm):
            ref_run_once(exe, qasm, env)
        for _ in range(steps):
            t, _, tail = ref_run_once(exe, qasm, env)
            if t is not None:
                times.append(t)
        kind = "reference"
        detail = (f"reference's own CUDA build, backend `{label}` (best of: {'; '.join(notes)}), sm_100, oracle/_ref, on GPUs 0..{world - 1} "
                  f"of this box ({'single-process peer-copy mode' if world > 1 else 'one GPU'}); 'Time Cost' of {len(times)} full runs of "
                  f"{name} after {warm + 1} warm-up run(s); steps bounded by a {budget:.0f} s budget; the reference has no CPU implementation")
        out["config"]["reference_backend"] = label
        out["config"]["reference_class"] = "gpu"
        warm_used = warm + 1
    if not times:
        est, cores, n_used, detail = cpu_replay(text, n)
        est *= 2.0 ** (n - n_used)
        times = [est]
        kind = "port"
        detail = ("the reference cannot run this size here (" + "; ".join(notes or ["L > 31 or no binaries"]) +
                  "); oracle OpenMP replay on host cores instead, NOT comparable as a GPU baseline: " + detail)
        out["config"]["reference_class"] = "cpu_port"
        warm_used = 0
    sec = sum(times) / len(times)
    value = bytes_per_step / sec / 1e12
    out.update({"value": value, "steps": len(times), "warmup": warm_used, "ms_per_step": sec * 1e3, "circuit_time_ms": sec * 1e3,
                "cpu_baseline": {"value": value, "unit": "TB/s", "cores": cores, "kind": kind, "sample": detail},
                "e2e": {"value": value, "unit": "TB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0})
    print(json.dumps(out), flush=True)


# ---------------------------------------------------------------------------------------------------------------
def parity_block(api, world, rank, dist, torch, names):
    """Small instances through the SAME product path (partitioner, specialised kernels, swaps), shards gathered on rank 0 and
    compared with the oracle.  Returns the dict for the JSON line (rank 0) and whether everything matched (all ranks)."""
    import numpy as np
    C = load_circuits()
    res = {"circuits": [], "max_abs_err": 0.0, "ok": True, "tolerance": 1e-10,
           "checked_by": "oracle/oracle.py gate-by-gate replay on rank 0; amplitudes gathered from all ranks + dump text"}
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
        if world > 1:
            mine = torch.from_numpy(shard.view(np.float64).copy()).cuda()
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            phys = np.concatenate([p.cpu().numpy().view(np.complex128) for p in parts]) if rank == 0 else None
        else:
            phys = shard
        if rank == 0:
            from oracle import oracle as O
            logical = np.arange(1 << n, dtype=np.int64)
            pid = np.zeros_like(logical)
            for q in range(n):
                pid |= ((logical >> q) & 1) << pos[q]
            got = phys[pid]
            _, gates = O.parse_qasm(text)
            want = O.simulate(n, gates)
            err = float(np.max(np.abs(got - want)))
            same_dump, _ = O.compare_dumps(O.dump_state(want, n), dump)
            ok = err <= 1e-10 and bool(same_dump)
            res["circuits"].append({"name": name, "stages": info["stages"], "groups": info["groups"], "max_abs_err": err,
                                    "dump_equal": bool(same_dump), "ok": ok})
            res["max_abs_err"] = max(res["max_abs_err"], err)
            res["ok"] = res["ok"] and ok
        c.close()
    flag = torch.tensor([1 if res["ok"] else 0], device="cuda")
    if world > 1:
        dist.broadcast(flag, 0)
    return res, bool(int(flag.item()))


def run_ours(args, world, rank, local_rank):
    # one OpenMP team per rank: the product compiles its specialised kernels on these threads (and rank 0 runs the oracle)
    if world > 1:
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or world) // world))
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries the one JSON line
    import torch
    from hyquas_b200 import api
    import ctypes
    from hyquas_b200._lib import check, lib

    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version line on stdout when NCCL_DEBUG is set in the environment; stdout carries the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
            torch.cuda.set_device(local_rank)
            dist.barrier()
            api.init()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    torch.cuda.set_device(local_rank)
    api.init()
    name, n, text = workload(args, world)
    g = int(math.log2(world))
    L = n - g
    peak, peak_src = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def jit_stats():
        a, b, c_, d = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
        check(lib.hq_jit_stats(a, b, c_, d))
        return {"kernels_loaded": a.value, "compiled": b.value, "disk_hits": c_.value, "compile_cpu_seconds": round(d.value, 2)}

    # ---- parity on the product path, before anything is timed ----------------------------------------------
    fam = args.circuit
    # small instances of the workload family and of qaoa (needs an even qubit count), 22-24 qubits in all, 10+ local qubits
    parity_names = [f"{fam}_{22 + g}", f"qaoa_{22 + 2 * ((g + 1) // 2)}"] if not args.no_parity else []
    parity, parity_ok = (None, True)
    if parity_names:
        parity, parity_ok = parity_block(api, world, rank, dist, torch, parity_names)
        if not parity_ok:
            if rank == 0:
                print(json.dumps({"metric": "effective_hbm_tbps", "value": None, "parity": parity, "error": "parity mismatch"}), flush=True)
            sys.exit(3)

    t0 = time.perf_counter()
    c = api.Circuit.from_qasm(text)
    c.compile()
    first_compile_s = time.perf_counter() - t0
    jit_after_first = jit_stats()
    info = c.schedule_info()
    S = info["groups"]
    S_nom = nominal_sweeps(name, n, text)
    bytes_per_step = S_nom * 32.0 * (1 << L) * world
    c.prepare_state()

    # ---- value: K executions on the resident state, CUDA events on the launching stream -------------------
    # nvidia-smi needs a few hundred ms to deliver its first sample: start it before the warm-up so that short timed regions
    # (5 steps of 70-90 ms) are covered too; samples therefore span warm-up + timed region, the same kernels at the same load
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(3, args.warmup)):
        c.execute()
    barrier()
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
    groups, pos = [], 0
    for gi in ginfo:   # execute() reports one entry per LAUNCH; a per-chunk group has several: fold them back per group
        ms = sum(per_launch[pos:pos + gi["launches"]])
        pos += gi["launches"]
        groups.append({"backend": gi["backend"], "gates": gi["gates"], "blocks": gi["blocks"], "launches": gi["launches"],
                       "ms": round(ms, 3), "predicted_ms": round(gi["predicted_ms"], 3)})   # (both summed over the group's launches)
    share = {}
    for gr in groups:
        share[gr["backend"]] = share.get(gr["backend"], 0.0) + gr["ms"]
    dominant = max(share, key=share.get) if share else "tile"
    dom = [gr["ms"] / gr["launches"] * (1 << 0) for gr in groups if gr["backend"] == dominant and gr["launches"] == 1]
    mean_launch_ms = sum(dom) / max(1, len(dom))
    alg_bytes = 32.0 * (1 << L)
    achieved = alg_bytes / (mean_launch_ms * 1e-3) / 1e9 if dom else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get(f"{dominant}_kernel_dram_bytes_per_launch")
    spec = sum(1 for _ in ginfo)  # placeholder for readability below
    roofline = {"bound": "hbm", "kernel": ("hq_group_jit (specialised tile kernel)" if jit_after_first["kernels_loaded"] else "group_kernel")
                if dominant == "tile" else "dense_kernel", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "launch_ms_mean": mean_launch_ms, "launch_ms_min": min(dom) if dom else None,
                "launch_ms_max": max(dom) if dom else None, "time_share": share.get(dominant, 0.0) / max(1e-9, sum(share.values())),
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "one launch = one in-place sweep of the local state (full-state launches only; per-chunk launches under an "
                        "exchange are listed in `groups`); gate-heavy launches are FP64-bound (see fp64)"}
    v = ctypes.c_double()
    check(lib.hq_microbench_fp64(0, v)); fma_tf = v.value
    check(lib.hq_microbench_fp64(1, v)); mma_tf = v.value
    roofline["fp64"] = {"fma_tflops_measured": fma_tf, "dmma_tflops_measured": mma_tf,
                        "source": "hq_microbench_fp64, run live in this process (MEASURED_PEAKS.json has no FP64 figure)"}

    # ---- how much of the exchange is hidden (N > 1): same circuit with the per-chunk overlap groups switched off -------------
    overlap = None
    if world > 1 and not args.no_overlap_probe:
        c.release_state()           # one state at a time; the p2p mapping follows the live state
        os.environ["HQ_ENABLE_OVERLAP"] = "0"
        c2 = api.Circuit.from_qasm(text)
        c2.compile()
        os.environ.pop("HQ_ENABLE_OVERLAP", None)
        c2.prepare_state()
        for _ in range(2):
            c2.execute()
        barrier()
        off_ms = 0.0
        reps = max(2, min(args.steps, 5))
        for _ in range(reps):
            off_ms += c2.execute()[1]
        barrier()
        t = torch.tensor([off_ms / reps], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        off_ms = float(t.item())
        sw = ctypes.c_double()
        check(lib.hq_circuit_swap_alone_ms(c2._h, sw))   # warm-up pass
        barrier()
        check(lib.hq_circuit_swap_alone_ms(c2._h, sw))
        t = torch.tensor([sw.value], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        swap_ms = float(t.item())
        off_sweeps = c2.schedule_info()["groups"]
        c2.close()
        overlap = {"overlap_groups": sum(1 for gi in ginfo if gi["launches"] > 1), "time_overlap_on_ms": ms_per_step,
                   "time_overlap_off_ms": off_ms, "sweeps_overlap_off": off_sweeps, "swap_alone_ms": swap_ms,
                   "hidden_frac": (off_ms - ms_per_step) / swap_ms if swap_ms > 0 else None,
                   "definition": "(T_overlap_off - T_overlap_on) / T_swap_alone; swap_alone = the schedule's exchanges run back to back with no compute"}
    c.close()

    # ---- e2e: the public API from host inputs (QASM text) to host outputs (amplitude dump) -----------------
    e2e_steps = max(1, min(args.steps, 5))
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
        for k, vv in zip(parts, (tb - ta, tc - tb, td - tc, te - td)):
            parts[k] += vv * 1e3 / e2e_steps
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": bytes_per_step / e2e_s / 1e12, "unit": "TB/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
           "breakdown_ms": {k: round(vv, 2) for k, vv in parts.items()},
           "path": "QASM text -> hq_circuit_from_qasm -> compile (kernels from the JIT cache) -> run(alloc, init, execute) -> dump"}

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
                          "nominal_sweeps": S_nom, "stages": info["stages"], "bytes_per_step": bytes_per_step,
                          "numerator": "frozen per workload (nominal_sweeps x 32 B x 2^n), identical in both arms",
                          "state_bytes_per_gpu": 16 * (1 << L), "l2": "inputs (16 GiB state) larger than L2",
                          "tile_bits": int(os.environ.get("HQ_TILE_BITS", "12")),
                          "backend": os.environ.get("HQ_BACKEND", "mix")},
               "circuit_time_ms": ms_per_step, "sweeps_per_s": S / (ms_per_step * 1e-3),
               "value_executed_schedule": S * 32.0 * (1 << L) * world / (ms_per_step * 1e-3) / 1e12,
               "roofline": roofline, "groups": groups, "cpu_baseline": cpu, "e2e": e2e, "parity": parity,
               "jit": {"first_compile_wall_s": round(first_compile_s, 2), "after_first_compile": jit_after_first, "at_exit": jit_stats(),
                       "note": "specialised tile kernels: NVRTC for sm_100a at Circuit::compile(), cached in memory and on disk; "
                               "first_compile_wall_s is this process's first compile() of the workload (cold unless the disk cache was warm)"},
               "overlap": overlap,
               "gpu_launches": sum(gi["launches"] for gi in ginfo) * args.steps,
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
    ap.add_argument("--no-parity", action="store_true", help="skip the product-path parity block")
    ap.add_argument("--no-overlap-probe", action="store_true", help="skip the overlap-off comparison run (N > 1)")
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
