#!/bin/bash
# usage: tools/gpu_session_multi4.sh N
set -u
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
F='^W\|^\*\*\*\|OMP_NUM_THREADS\|^$'
echo "== worker parity"; HQ_OVERLAP_SLACK=1e9 timeout 900 $TR --master-port 29721 tests/gpu_multirank_worker.py qft_20 supremacy_22 qaoa_22 adder_20 quantum_volume_20 hidden_shift_20 basis_change_18 bv_20 2>&1 | grep -v "$F" | tail -10 | tee gpurun_out/multi_parity_$N.log
echo "== swap bench"; timeout 600 $TR --master-port 29722 tools/swap_bench.py 30 2>&1 | grep -v "$F" | tail -2 | tee gpurun_out/swap_bench_p2p_$N.json
echo "== suite"; timeout 1200 $TR --master-port 29723 tools/run_suite.py ${SUITE:-qaoa_34 supremacy_32 quantum_volume_32 qft_32} 2>&1 | grep -v "$F" | tee gpurun_out/suite_${N}gpu.jsonl
echo "== suite no-overlap"; HQ_ENABLE_OVERLAP=0 timeout 900 $TR --master-port 29724 tools/run_suite.py ${SUITE_NO:-qaoa_34 supremacy_32} 2>&1 | grep -v "$F" | tee gpurun_out/suite_${N}gpu_nooverlap.jsonl
echo "== bench"; timeout 900 $TR --master-port 29725 bench.py --gpus $N --steps 5 --warmup 3 2>&1 | grep -v "$F" | tail -2 | tee gpurun_out/bench_multi_$N.json | cut -c1-400
