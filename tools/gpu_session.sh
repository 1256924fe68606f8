#!/bin/bash
# r02 session 3: two-worker / three-buffer skeleton (six landed-barriers) -- parity, microbench, bench, suite, ncu
set -u
O=gpurun_out/s3; mkdir -p $O
echo "== pytest gpu (JIT on)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
if ! grep -q " passed" $O/pytest_gpu.log || grep -q "failed" $O/pytest_gpu.log; then echo "PARITY FAILED - stopping"; exit 1; fi
CASES=sweep_1gate,h_x64_4q,u3_x64_4q,cu1fan_x64_4q,mix5_x64_4q,h_x96_12q,mix5_x96_12q,u3_x96_12q,h_x256_4q,sup5_x96_12q,mix5_x256_12q,h_x16_4q
echo "== microbench JIT on"; timeout 600 python tools/microbench.py --qubits 30 --only $CASES --out $O/microbench_jit.json 2>&1 | tee $O/microbench_jit.log | grep -E "_x|sweep"
for b in group mix; do
echo "== bench $b JIT on"; HQ_BACKEND=$b timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > $O/bench_${b}_jit.json 2> $O/bench_${b}_jit.err; cut -c1-330 $O/bench_${b}_jit.json; tail -3 $O/bench_${b}_jit.err
done
echo "== suite 1 gpu"; HQ_SUITE_PER_GROUP=1 timeout 900 python tools/run_suite.py qft_28 qft_30 qaoa_30 quantum_volume_30 bv_30 hidden_shift_30 adder_30 basis_change_28 2>/dev/null | tee $O/suite_1gpu.jsonl | cut -c1-330
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-900 $O/bench_reference.json
echo "== ncu full: supremacy_30 group backend, 2 jit launches"
HQ_BACKEND=group timeout 900 ncu --set full --clock-control none --import-source on -k regex:hq_group_jit -s 24 -c 2 -f -o $O/prof_sup_group_jit \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-parity > $O/ncu_sup_group.log 2>&1; tail -2 $O/ncu_sup_group.log
ls -la $O | tail -14
