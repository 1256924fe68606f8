#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu (default lib, RBITS=3)"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_r3.log
for sfx in "" "_r4"; do
  echo "== microbench lib$sfx"; HQ_LIB_SUFFIX=$sfx timeout 600 python tools/microbench.py --qubits 30 --out gpurun_out/microbench$sfx.json 2>&1 | tee gpurun_out/microbench$sfx.log
  echo "== bench lib$sfx"; HQ_LIB_SUFFIX=$sfx timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tee gpurun_out/bench$sfx.json
done
