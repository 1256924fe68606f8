#!/bin/bash
# Multi-GPU acceptance run (one `gpurun --gpus N` call): product-path parity, launcher mode, bench both arms, suite with overlap probe.
#   gpurun --gpus 8 --timeout 1500 -- "LOG2N=3 bash tools/gpu_check_multi.sh 8"     (outputs under gpurun_out/m8/)
set -u
N=${1:-2}
O=gpurun_out/m$N; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi --query-gpu=name --format=csv,noheader | head -$N | tr '\n' ';'; nproc
if [ -z "${SKIP_PYTEST:-}" ]; then echo "== multi-GPU parity (pytest, world $N)"; timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "$N" > $O/pytest_multi.log 2>&1; tail -3 $O/pytest_multi.log; grep "multi-gpu x" $O/pytest_multi.log | head -12; fi
echo "== launcher mode: hyquas_main on all GPUs vs golden"
( timeout 600 ./hyquas_b200/hyquas_main tests/golden/qft_28.qasm > $O/main_qft28.log 2> $O/main_qft28.err; grep -Ev "Logger|CLUSTER" $O/main_qft28.log | diff -q - tests/golden/qft_28.log && echo "qft_28 golden: identical"; grep -E "Time Cost|Total Groups" $O/main_qft28.log | head -4; tail -2 $O/main_qft28.err )
echo "== bench ours N=$N"; timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_ours.json 2> $O/bench_ours.err; cut -c1-250 $O/bench_ours.json; tail -2 $O/bench_ours.err
python - <<P
import json
try:
    d=json.loads(open("$O/bench_ours.json").read()); print("ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], "parity", d["parity"]["ok"], d["parity"]["max_abs_err"], "overlap", d["overlap"]); print([ (g["backend"][0], g["gates"], g["launches"], g["ms"]) for g in d["groups"]])
except Exception as e: print("no json", e)
P
if [ -z "${SKIP_REF:-}" ]; then echo "== bench reference N=$N"; timeout 900 $TR bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; cut -c1-1200 $O/bench_ref.json; tail -2 $O/bench_ref.err; fi
echo "== suite N=$N"; HQ_SUITE_OVERLAP_PROBE=1 timeout 1200 $TR tools/run_suite.py ${SUITE:-supremacy_$((30 + ${LOG2N:-1})) qaoa_$((30 + ${LOG2N:-1})) qft_$((30 + ${LOG2N:-1}))} 2> $O/suite.err | tee $O/suite.jsonl | cut -c1-700; tail -2 $O/suite.err
