#!/bin/bash
# ncu full capture of the gate-group kernel on the bench workload at 28 qubits (one gpurun call).
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:group_kernel -s ${SKIP:-6} -c ${COUNT:-2} -f -o gpurun_out/prof_group \
    python bench.py --steps 1 --warmup 1 --no-cpu --qubits ${QUBITS:-28} ${BENCH_ARGS:-} > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
