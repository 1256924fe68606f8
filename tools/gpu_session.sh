#!/bin/bash
# r02 session 2: two-worker / three-buffer skeleton -- parity, microbench, bench; reference parameter files + mix build
set -u
O=gpurun_out/s2; mkdir -p $O
echo "== pytest gpu (JIT on)"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
CASES=sweep_1gate,h_x64_4q,u3_x64_4q,t_x64_4q,cu1fan_x64_4q,czfan_x64_4q,mix5_x64_4q,h_x96_12q,mix5_x96_12q,u3_x96_12q,h_x256_4q,sup5_x64_4q,sup5_x96_12q,mix5_x256_12q,h_x16_4q
echo "== microbench JIT on"; timeout 900 python tools/microbench.py --qubits 30 --only $CASES --out $O/microbench_jit.json 2>&1 | tee $O/microbench_jit.log | grep -E "_x|sweep"
for b in group mix; do
echo "== bench $b JIT on"; HQ_BACKEND=$b timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > $O/bench_${b}_jit.json 2> $O/bench_${b}_jit.err; cut -c1-330 $O/bench_${b}_jit.json; tail -3 $O/bench_${b}_jit.err
done
echo "== suite 1 gpu"; HQ_SUITE_PER_GROUP=1 timeout 900 python tools/run_suite.py qft_28 qft_30 qaoa_30 quantum_volume_30 bv_30 hidden_shift_30 adder_30 basis_change_28 2>/dev/null | tee $O/suite_1gpu.jsonl | cut -c1-400
echo "== reference parameter files (process 28 30) + mix build"
( cd oracle/_ref/run && CUDA_VISIBLE_DEVICES=0 timeout 900 ../hyquas_ref_process 30 > ../../../$O/ref_process.log 2>&1; tail -2 ../../../$O/ref_process.log; cp ../evaluator-preprocess/parameter-files/*.out ../../../$O/ 2>/dev/null
  python -c "
import sys; sys.path.insert(0,'../../..')
from hyquas_b200 import circuits
open('/tmp/sup30.qasm','w').write(circuits.generate('supremacy_30'))"
  for b in 1 3 4p; do echo "-- ref b$b"; CUDA_VISIBLE_DEVICES=0 timeout 600 ../hyquas_ref_b$b /tmp/sup30.qasm 2>&1 | grep -E "Time Cost|Total Groups|not find|rror" | head -5; done )
echo "== ncu full: supremacy_30 group backend, 2 jit launches"
HQ_BACKEND=group timeout 1200 ncu --set full --clock-control none --import-source on -k regex:hq_group_jit -s 24 -c 2 -f -o $O/prof_sup_group_jit \
    python bench.py --steps 1 --warmup 3 --no-cpu > $O/ncu_sup_group.log 2>&1; tail -2 $O/ncu_sup_group.log
ls -la $O | tail -14
