#!/bin/bash
# overlap experiments on N GPUs: off / whole-group deferral / split deferral, and co-resident CTA count
set -u
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
F='^W\|^\*\*\*\|OMP_NUM_THREADS\|^$\|NCCL version'
CIRC=${CIRC:-"qaoa_32 supremacy_31 quantum_volume_31 hidden_shift_32 adder_32 bv_32"}
port=29800
for mode in off groups split; do
  port=$((port+1))
  echo "== mode $mode"; HQ_OVERLAP_MODE=$mode timeout 900 $TR --master-port $port tools/run_suite.py $CIRC 2>&1 | grep -v "$F" | tee gpurun_out/overlap_${N}gpu_$mode.jsonl | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['circuit'], 'ms', d['time_ms'], 'sweeps', d['sweeps'], 'overlap', d['overlap_groups'], 'pred', d['predicted_ms'], 'ok', d['ok'])
"
done
for ctas in 48 148; do
  port=$((port+1))
  echo "== split, co-resident ctas $ctas"; HQ_SWAP_CTAS_OVERLAP=$ctas HQ_OVERLAP_MODE=split timeout 900 $TR --master-port $port tools/run_suite.py $CIRC 2>&1 | grep -v "$F" | tee gpurun_out/overlap_${N}gpu_split_ctas$ctas.jsonl | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(d['circuit'], 'ms', d['time_ms'], 'sweeps', d['sweeps'], 'overlap', d['overlap_groups'])
"
done
echo "== swap bench"; timeout 600 $TR --master-port 29850 tools/swap_bench.py 30 2>&1 | grep -v "$F" | tail -1 | cut -c1-600
