#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest dense"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k dense 2>&1 | tail -3
echo "== microbench"; timeout 900 python tools/microbench.py --qubits 30 --out gpurun_out/microbench_s6.json 2>&1 | grep -E "dense|fp64|dmma" | tee gpurun_out/microbench_s6.log
