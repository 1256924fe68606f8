#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest dense"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for db in 1 0; do
echo "== microbench dense DB=$db"; HQ_DENSE_DB=$db timeout 900 python tools/microbench.py --qubits 30 --out gpurun_out/microbench_s12_db$db.json 2>&1 | grep -E "^dense"
echo "== bench DB=$db"; HQ_DENSE_DB=$db timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tee gpurun_out/bench_s12_db$db.json | cut -c1-250
done
echo "== suite"; timeout 900 python tools/run_suite.py qaoa_30 quantum_volume_30 basis_change_28 hidden_shift_28 bv_28 2>&1 | tee gpurun_out/suite_s12.jsonl | cut -c1-330
