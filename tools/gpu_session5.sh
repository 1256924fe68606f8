#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_s5.log
echo "== microbench"; timeout 900 python tools/microbench.py --qubits 30 --out gpurun_out/microbench_s5.json 2>&1 | grep -E "dense|fp64" | tee gpurun_out/microbench_s5.log
for mode in group blas mix; do
  echo "== bench supremacy_30 $mode"; HQ_BACKEND=$mode timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tee gpurun_out/bench_s5_sup30_$mode.json | cut -c1-400
  echo "== bench quantum_volume_30 $mode"; HQ_BACKEND=$mode timeout 600 python bench.py --circuit quantum_volume --steps 3 --warmup 3 --no-cpu 2>&1 | tee gpurun_out/bench_s5_qv30_$mode.json | cut -c1-400
done
for circ in qft qaoa; do
  for mode in group mix; do
  echo "== bench $circ $mode"; HQ_BACKEND=$mode timeout 600 python bench.py --circuit $circ --steps 3 --warmup 3 --no-cpu 2>&1 | tee gpurun_out/bench_s5_${circ}30_$mode.json | cut -c1-400
  done
done
echo "== reference on quantum_volume_30"; timeout 900 python bench.py --impl reference --circuit quantum_volume --steps 2 --warmup 1 2>&1 | tee gpurun_out/bench_s5_ref_qv30.json | cut -c1-600
