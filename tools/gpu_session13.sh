#!/bin/bash
# session 13: butterfly class + brx.idx dispatch + header read-ahead in the tile kernel
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu (single GPU)"; timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_s13.log
echo "== microbench"; timeout 900 python tools/microbench.py --qubits 30 --out gpurun_out/microbench_s13.json 2>&1 | tee gpurun_out/microbench_s13.log | grep -E "_x|sweep"
echo "== microbench, butterflies off"; HQ_NO_BUTTERFLY=1 timeout 600 python tools/microbench.py --qubits 30 --only h_x64_4q,h_x96_12q,sup5_x64_4q,sup5_x96_12q --out gpurun_out/microbench_s13_nobf.json 2>&1 | grep -E "_x"
for b in group mix; do
echo "== bench $b"; HQ_BACKEND=$b timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tee gpurun_out/bench_s13_$b.json | cut -c1-400
done
echo "== ncu full: h_x96_12q (3 rounds of H)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:group_kernel -s 3 -c 1 -f -o gpurun_out/prof_h96 \
    python tools/microbench.py --qubits 30 --only h_x96_12q > gpurun_out/ncu_h96.log 2>&1
echo "== ncu full: supremacy_30 group backend, 2 launches"
HQ_BACKEND=group timeout 1200 ncu --set full --clock-control none --import-source on -k regex:group_kernel -s 32 -c 2 -f -o gpurun_out/prof_sup_group \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_sup_group.log 2>&1
ls -la gpurun_out | tail -8
