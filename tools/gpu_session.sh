#!/bin/bash
# r02 session 5: tile geometry vs sweep speed (run length, spread of the tile over the state)
set -u
O=gpurun_out/s5; mkdir -p $O
G=sweep_lo7,sweep_mid7,sweep_hi7,sweep_spread7,sup5_x42_lo7,sup5_x42_mid7,sup5_x42_hi7,sup5_x42_spread7,sup5_x84_lo7,sup5_x84_mid7,sup5_x84_hi7,sup5_x84_spread7
for pb in 5 6 4; do
echo "== HQ_PINNED_BITS=$pb"; HQ_PINNED_BITS=$pb timeout 600 python tools/microbench.py --qubits 30 --only $G --out $O/microbench_geom_pb$pb.json 2>&1 | grep -E "sweep_|sup5" | cut -c1-100
done
echo "== JIT off, pinned 5"; HQ_JIT=0 timeout 600 python tools/microbench.py --qubits 30 --only sweep_lo7,sweep_hi7,sweep_spread7 2>&1 | grep -E "sweep_" | cut -c1-100
