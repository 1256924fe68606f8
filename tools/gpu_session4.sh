#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest dense"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k dense 2>&1 | tail -8 | tee gpurun_out/pytest_dense.log
echo "== microbench"; timeout 900 python tools/microbench.py --qubits 30 --out gpurun_out/microbench_s4.json 2>&1 | tee gpurun_out/microbench_s4.log
