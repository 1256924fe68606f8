#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_s10.log
echo "== microbench"; timeout 900 python tools/microbench.py --qubits 30 --out gpurun_out/microbench_s10.json 2>&1 | tee gpurun_out/microbench_s10.log | grep -E "_x|dmma"
echo "== suite qft"; timeout 600 python tools/run_suite.py qft_28 qft_30 adder_28 supremacy_30 2>&1 | tee gpurun_out/suite_s10.jsonl | cut -c1-420
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tee gpurun_out/bench_s10.json | cut -c1-300
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_s10.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches_s10.log 2>&1
echo "== ncu full dense"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:dense_kernel -s 16 -c 3 -f -o gpurun_out/prof_dense \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_dense.log 2>&1
echo "== ncu full tile"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:group_kernel -s 4 -c 2 -f -o gpurun_out/prof_tile2 \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_tile2.log 2>&1
ls -la gpurun_out | tail -5
