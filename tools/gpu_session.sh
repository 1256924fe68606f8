#!/bin/bash
# One GPU-box session: GPU parity tests, bench (ours + reference), ncu launch list, ncu full capture of the top kernel.
# Everything lands in gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/box.txt 2>&1
nproc >> gpurun_out/box.txt; free -g >> gpurun_out/box.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench ours"; timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_ours.err | tee gpurun_out/bench_ours.json
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches_bench.log 2>&1
echo "== ncu full"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:group_kernel -s 12 -c 3 -f -o gpurun_out/prof_group \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out
