#!/bin/bash
# r02 session 1: specialised (JIT) tile kernels -- parity, microbench A/B, bench A/B, one ncu capture
set -u
O=gpurun_out/s1; mkdir -p $O
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | tee $O/gpu.txt; nproc | tee -a $O/gpu.txt
echo "== pytest gpu (JIT on)"; HQ_JIT_VERBOSE=1 timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log; grep -c "hq jit. compiled" $O/pytest_gpu.log
CASES=sweep_1gate,h_x64_4q,rx_x64_4q,u3_x64_4q,t_x64_4q,rz_x64_4q,x_x64_4q,cz_x64_4q,cnot_x64_4q,cu1fan_x64_4q,czfan_x64_4q,mix5_x64_4q,h_x96_12q,mix5_x96_12q,u3_x96_12q,h_x256_4q,sup5_x64_4q,sup5_x96_12q,mix5_x256_12q,t_x96_12q
echo "== microbench JIT on"; timeout 900 python tools/microbench.py --qubits 30 --only $CASES --out $O/microbench_jit.json 2>&1 | tee $O/microbench_jit.log | grep -E "_x|sweep"
echo "== microbench JIT off"; HQ_JIT=0 timeout 900 python tools/microbench.py --qubits 30 --only h_x64_4q,u3_x64_4q,sup5_x96_12q,mix5_x96_12q --out $O/microbench_nojit.json 2>&1 | grep -E "_x|sweep"
for b in group mix; do
echo "== bench $b JIT on"; HQ_JIT_VERBOSE=1 HQ_BACKEND=$b timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > $O/bench_${b}_jit.json 2> $O/bench_${b}_jit.err; cut -c1-600 $O/bench_${b}_jit.json; grep "hq jit. compiled" $O/bench_${b}_jit.err | head -30
done
echo "== bench group JIT off"; HQ_JIT=0 HQ_BACKEND=group timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > $O/bench_group_nojit.json 2>/dev/null; cut -c1-300 $O/bench_group_nojit.json
echo "== ncu full: supremacy_30 group backend, 2 jit launches"
HQ_BACKEND=group timeout 1200 ncu --set full --clock-control none --import-source on -k regex:hq_group_jit -s 24 -c 2 -f -o $O/prof_sup_group_jit \
    python bench.py --steps 1 --warmup 3 --no-cpu > $O/ncu_sup_group.log 2>&1; tail -2 $O/ncu_sup_group.log
ls -la $O | tail -12
