#!/bin/bash
# r02 session 7 (1 GPU): full GPU test-suite, default bench (both arms), ncu launch list + full captures (tile + dense), suite incl. QV_32
set -u
O=gpurun_out/s7; mkdir -p $O
echo "== pytest -m gpu (all)"; timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
if grep -q "failed" $O/pytest_gpu.log; then grep -E "^E |FAILED" $O/pytest_gpu.log | head -20; fi
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-400 $O/bench_reference.json
echo "== bench ours (default flags)"; timeout 900 python bench.py > $O/bench_ours.json 2> $O/bench_ours.err; tail -2 $O/bench_ours.err
python - <<P
import json
d=json.loads([l for l in open("$O/bench_ours.json") if l.startswith("{")][0])
print("ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["breakdown_ms"], "roofline", d["roofline"]["frac"], d["roofline"]["kernel"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], "clocks", d["clocks"], "jit", d["jit"]["first_compile_wall_s"])
print([(g["gates"], g["ms"], g["predicted_ms"]) for g in d["groups"]])
P
echo "== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-parity > $O/launches_bench.log 2>&1; tail -1 $O/launches_bench.log | cut -c1-200
echo "== ncu full: supremacy_30, 3 specialised launches"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hq_group_jit -s 20 -c 3 -f -o $O/prof_sup_jit python bench.py --steps 1 --warmup 3 --no-cpu --no-parity > $O/ncu_sup.log 2>&1; tail -1 $O/ncu_sup.log | cut -c1-200
echo "== ncu full: quantum_volume_30, 3 dense launches"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_kernel -s 12 -c 3 -f -o $O/prof_qv_dense python bench.py --circuit quantum_volume --qubits 30 --steps 1 --warmup 3 --no-cpu --no-parity > $O/ncu_qv.log 2>&1; tail -1 $O/ncu_qv.log | cut -c1-200
echo "== suite 1 gpu"; HQ_SUITE_PER_GROUP=1 timeout 1200 python tools/run_suite.py qft_28 bv_28 hidden_shift_28 supremacy_30 qft_30 qaoa_30 quantum_volume_30 bv_30 hidden_shift_30 adder_30 basis_change_28 quantum_volume_32 2>/dev/null | tee $O/suite_1gpu.jsonl | cut -c1-230
for b in group blas; do echo "== supremacy_30 backend $b"; HQ_BACKEND=$b timeout 600 python tools/run_suite.py supremacy_30 2>/dev/null | tee -a $O/suite_backends.jsonl | cut -c1-230; done
