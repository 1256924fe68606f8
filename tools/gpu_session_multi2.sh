#!/bin/bash
# usage: tools/gpu_session_multi2.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
F='^W\|^\*\*\*\|OMP_NUM_THREADS\|^$'
echo "== pytest multi (p2p)"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s 2>&1 | grep -v "$F" | tail -16 | tee gpurun_out/pytest_multi2_$N.log
echo "== forced overlap (p2p)"; HQ_OVERLAP_SLACK=1e9 timeout 600 $TR --master-port 29711 tests/gpu_multirank_worker.py qft_20 supremacy_22 qaoa_22 adder_20 quantum_volume_20 2>&1 | grep -v "$F" | tail -8 | tee gpurun_out/multi2_forced_overlap_$N.log
echo "== nccl transport"; HQ_SWAP=nccl HQ_OVERLAP_SLACK=1e9 timeout 600 $TR --master-port 29713 tests/gpu_multirank_worker.py qft_20 supremacy_22 qaoa_22 2>&1 | grep -v "$F" | tail -6 | tee gpurun_out/multi2_nccl_$N.log
echo "== swap bench p2p"; timeout 600 $TR --master-port 29714 tools/swap_bench.py 30 2>&1 | grep -v "$F" | tail -3 | tee gpurun_out/swap_bench_p2p_$N.json
echo "== swap bench nccl"; HQ_SWAP=nccl timeout 600 $TR --master-port 29715 tools/swap_bench.py 30 2>&1 | grep -v "$F" | tail -3 | tee gpurun_out/swap_bench_nccl_$N.json
echo "== suite"; timeout 900 $TR --master-port 29716 tools/run_suite.py supremacy_31 qaoa_31 quantum_volume_31 qft_31 2>&1 | grep -v "$F" | tee gpurun_out/suite_${N}gpu.jsonl
echo "== suite no-overlap"; HQ_ENABLE_OVERLAP=0 timeout 900 $TR --master-port 29717 tools/run_suite.py supremacy_31 qaoa_31 2>&1 | grep -v "$F" | tee gpurun_out/suite_${N}gpu_nooverlap.jsonl
