#!/bin/bash
# Sweep kernel configurations on the bench workload (one gpurun call).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for cfg in "12 0" "12 1" "11 0" "11 1" "10 0"; do
  set -- $cfg
  echo "== HQ_TILE_BITS=$1 HQ_RELAXED_REGS=$2"
  HQ_TILE_BITS=$1 HQ_RELAXED_REGS=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu ${BENCH_ARGS:-} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
r=d['roofline']
print('circuit_ms',round(d['ms_per_step'],2),'sweeps',d['config']['sweeps'],'launch_ms min/mean/max',round(r['launch_ms_min'],2),round(r['launch_ms_mean'],2),round(r['launch_ms_max'],2),'frac',round(r['frac'],3),'e2e_ms',round(d['e2e']['ms_per_step'],1))
"
done
