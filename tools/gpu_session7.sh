#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_s7.log
echo "== microbench"; timeout 900 python tools/microbench.py --qubits 30 --out gpurun_out/microbench_s7.json 2>&1 | tee gpurun_out/microbench_s7.log | grep -E "dense|_x"
echo "== bench mix"; timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tee gpurun_out/bench_s7_mix.json | cut -c1-300
echo "== suite 1 GPU"; timeout 1200 python tools/run_suite.py qft_28 bv_28 hidden_shift_28 adder_28 supremacy_30 qaoa_30 basis_change_28 quantum_volume_32 2>&1 | tee gpurun_out/suite_1gpu.jsonl
