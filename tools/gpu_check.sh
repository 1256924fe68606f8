#!/bin/bash
# One gpurun call: GPU tests, a bench line, the ncu launch list and one full capture of the dominant kernel.
#   gpurun --timeout 1500 -- 'bash tools/gpu_check.sh'
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.txt
echo "== bench"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
grep -c group_kernel gpurun_out/launches.csv
echo "== ncu full capture (28 qubits)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:group_kernel -s 4 -c 3 -f -o gpurun_out/prof_group \
    python bench.py --steps 1 --warmup 1 --no-cpu --qubits 28 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
