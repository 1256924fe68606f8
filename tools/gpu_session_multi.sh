#!/bin/bash
# usage: tools/gpu_session_multi.sh N   (run under gpurun --gpus N)
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/multi_gpus.txt
nvidia-smi topo -m >> gpurun_out/multi_gpus.txt 2>&1
echo "== pytest multi"; timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s 2>&1 | tail -25 | tee gpurun_out/pytest_multi_$N.log
echo "== forced overlap"; HQ_OVERLAP_SLACK=1e9 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 tests/gpu_multirank_worker.py qft_20 supremacy_22 qaoa_22 adder_20 quantum_volume_20 2>&1 | grep -v "^W\|^\*\*\*" | tail -12 | tee gpurun_out/multi_forced_overlap_$N.log
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus $N --steps 5 --warmup 3 2>&1 | grep -v "^W\|^\*\*\*" | tail -5 | tee gpurun_out/bench_multi_$N.json
