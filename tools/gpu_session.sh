#!/bin/bash
# r02 session 9 (1 GPU): pinned low bits of every tile (HBM run length 512 / 256 / 128 B) vs number of sweeps
set -u
O=gpurun_out/s9; mkdir -p $O
for pb in 5 4 3; do
echo "== HQ_PINNED_BITS=$pb"; HQ_PINNED_BITS=$pb HQ_SUITE_PER_GROUP=1 timeout 600 python tools/run_suite.py supremacy_30 qaoa_30 qft_30 basis_change_28 hidden_shift_30 2>/dev/null | tee $O/suite_pb$pb.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['circuit'], d['sweeps'], d['time_ms'], d['ok'], [round(x,2) for x in d['per_group']['launch_ms']][:12])
"
done
G=sweep_lo7,sweep_hi7,sweep_spread7,sup5_x42_hi7,sup5_x42_spread7
for pb in 5 3; do echo "== microbench geometry, pinned $pb (tile = pinned bits + 7 chosen + fill)"; HQ_PINNED_BITS=$pb timeout 300 python tools/microbench.py --qubits 30 --only $G 2>&1 | grep -E "sweep_|sup5" | cut -c1-90; done
