#!/bin/bash
# r02 session 8 (1 GPU): plan-identity JIT cache (e2e compile time), quantum_volume_30 A/B (fusion-aware evaluator, tile-only)
set -u
O=gpurun_out/s8; mkdir -p $O
echo "== pytest gpu (parity file)"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
echo "== bench ours"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench_ours.json 2> $O/bench_ours.err; tail -2 $O/bench_ours.err
python - <<P
import json
d=json.loads([l for l in open("$O/bench_ours.json") if l.startswith("{")][0])
print("ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["breakdown_ms"], "clocks", d["clocks"], "jit", d["jit"])
P
echo "== second process: disk cache warm"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-parity > $O/bench_warm.json 2>/dev/null
python - <<P
import json
d=json.loads([l for l in open("$O/bench_warm.json") if l.startswith("{")][0])
print("ms", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["breakdown_ms"], "jit", d["jit"])
P
echo "== quantum_volume_30: mix / mix fusion-aware / tile-only"
timeout 300 python tools/run_suite.py quantum_volume_30 2>/dev/null | tee $O/qv30_mix.jsonl | cut -c1-260
HQ_EVAL_FUSION=1 timeout 300 python tools/run_suite.py quantum_volume_30 2>/dev/null | tee $O/qv30_mix_fusionaware.jsonl | cut -c1-260
HQ_BACKEND=group timeout 300 python tools/run_suite.py quantum_volume_30 2>/dev/null | tee $O/qv30_group.jsonl | cut -c1-260
HQ_BACKEND=group HQ_JIT_NO_FUSE=1 timeout 300 python tools/run_suite.py quantum_volume_30 2>/dev/null | tee $O/qv30_group_nofuse.jsonl | cut -c1-260
