#!/bin/bash
# One-GPU acceptance run (one `gpurun` call): full GPU test-suite, smoke(), bench both arms with driver-style flags, CLI vs golden.
#   gpurun --timeout 1500 -- bash tools/gpu_check.sh        (outputs under gpurun_out/check/)
set -u
O=gpurun_out/check; mkdir -p $O
echo "== pytest -m gpu (all)"; timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
if grep -q "failed" $O/pytest_gpu.log; then grep -E "^E |FAILED" $O/pytest_gpu.log | head -20; fi
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench reference (driver-style flags)"; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-300 $O/bench_reference.json
echo "== bench ours (driver-style flags)"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_ours.json 2> $O/bench_ours.err; tail -2 $O/bench_ours.err
python - <<P
import json
d=json.loads([l for l in open("$O/bench_ours.json") if l.startswith("{")][0])
print("ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["breakdown_ms"], "roofline", d["roofline"]["frac"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"][:80], "clocks", d["clocks"], "launches", d["gpu_launches"], "parity", d["parity"]["ok"])
print([(g["gates"], g["ms"], g["predicted_ms"]) for g in d["groups"]])
P
echo "== hyquas_main, single GPU, golden"; ./hyquas_b200/hyquas_main tests/golden/bv_28.qasm 2>/dev/null | grep -v Logger | diff -q - tests/golden/bv_28.log && echo "bv_28 golden identical"
echo "== cut search A/B through the CLI (supremacy_30, plain vs searched cut; r02_s13)"
python -c "
import sys; sys.path.insert(0, '.')
from hyquas_b200 import circuits as C
open('$O/s30.qasm', 'w').write(C.supremacy(30))"
for v in 1 32; do for i in 1 2; do
  HQ_NUM_GPUS=1 HQ_CUT_VARIANTS=$v ./hyquas_b200/hyquas_main $O/s30.qasm > $O/cli_v${v}_$i.out 2> $O/cli_v${v}_$i.err
  echo "HQ_CUT_VARIANTS=$v run $i: $(grep 'Time Cost\|Total Groups' $O/cli_v${v}_$i.out | tr '\n' ' ') dump $(grep -E '^[0-9]+ [0-9.]+: ' $O/cli_v${v}_$i.out | sha256sum | cut -c1-16)"
done; done
