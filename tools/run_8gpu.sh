#!/bin/bash
# One 8-GPU session: the 36-qubit weak-scaling suite (configs[4]), the 8-GPU bench line, qaoa_34 strong scaling at 8/4/2 GPUs
# (configs[3]) with and without overlap, multi-GPU parity, swap bandwidth.  Every step has its own timeout; outputs in $1.
OUT=${1:-gpurun_out/s18}
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=name,memory.total --format=csv > $OUT/gpus.csv 2>&1
timeout 200 $TR --nproc-per-node 8 --master-port 29801 tools/run_suite.py bv_36 hidden_shift_36 adder_36 basis_change_36 > $OUT/suite_36q_8gpu.jsonl 2> $OUT/suite_36q_8gpu.err
timeout 120 $TR --nproc-per-node 8 --master-port 29802 bench.py --gpus 8 --steps 5 --warmup 3 > $OUT/bench_8gpu.json 2> $OUT/bench_8gpu.err
timeout 90 $TR --nproc-per-node 8 --master-port 29803 tools/run_suite.py qaoa_34 supremacy_33 qft_33 > $OUT/suite_8gpu.jsonl 2> $OUT/suite_8gpu.err
HQ_ENABLE_OVERLAP=0 timeout 90 $TR --nproc-per-node 8 --master-port 29804 tools/run_suite.py qaoa_34 supremacy_33 > $OUT/suite_8gpu_nooverlap.jsonl 2> $OUT/suite_8gpu_nooverlap.err
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "4 or 8" > $OUT/pytest_multi.log 2>&1
timeout 90 $TR --nproc-per-node 4 --master-port 29805 tools/run_suite.py qaoa_34 > $OUT/suite_qaoa34_4gpu.jsonl 2> $OUT/suite_qaoa34_4gpu.err
timeout 90 $TR --nproc-per-node 2 --master-port 29806 tools/run_suite.py qaoa_34 > $OUT/suite_qaoa34_2gpu.jsonl 2> $OUT/suite_qaoa34_2gpu.err
timeout 60 $TR --nproc-per-node 8 --master-port 29807 tools/swap_bench.py 30 > $OUT/swap_bench_8gpu.json 2>&1
echo finished > $OUT/done.txt
