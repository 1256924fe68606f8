#!/bin/bash
# r02 session 6: whole-worker TMA issue, group rebalancing, L2 prefetch (really on this time), suite, reference-build parity at 30 qubits
set -u
O=gpurun_out/s6; mkdir -p $O
echo "== pytest gpu"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
if grep -q "failed" $O/pytest_gpu.log; then echo "PARITY FAILED - stopping"; grep -E "^E " $O/pytest_gpu.log | head -20; exit 1; fi
G=sweep_lo7,sweep_hi7,sup5_x42_lo7,sup5_x42_mid7,sup5_x42_hi7,sup5_x42_spread7,sup5_x84_lo7,sup5_x84_hi7,h_x256_4q,u3_x64_4q
echo "== microbench"; timeout 600 python tools/microbench.py --qubits 30 --only $G --out $O/microbench.json 2>&1 | grep -E "sweep_|sup5|_x" | cut -c1-100
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-parity > $O/bench_$name.json 2> $O/bench_$name.err
  python - <<P
import json
try:
    d=json.loads(open("$O/bench_$name.json").read())
    print("$name", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["ms_per_step"],1), [g["ms"] for g in d["groups"]], d["jit"]["first_compile_wall_s"])
except Exception as e: print("$name failed", e, open("$O/bench_$name.err").read()[-300:])
P
}
run group HQ_BACKEND=group
run group_norebalance HQ_BACKEND=group HQ_REBALANCE=0
run group_pf6 HQ_BACKEND=group HQ_JIT_L2_PREFETCH=6
run mix HQ_BACKEND=mix
echo "== suite 1 gpu"; HQ_SUITE_PER_GROUP=1 timeout 900 python tools/run_suite.py qft_28 qft_30 qaoa_30 quantum_volume_30 bv_30 hidden_shift_30 adder_30 basis_change_28 2>/dev/null | tee $O/suite_1gpu.jsonl | cut -c1-260
