#!/usr/bin/env python
"""Turns a tools/microbench.py JSON into an evaluator parameter file ($HYQUAS_PARAM_FILE).

    python tools/microbench.py --qubits 30 --out mb.json      (on the GPU box)
    python tools/calibrate.py mb.json > b200.params
    HYQUAS_PARAM_FILE=b200.params ./hyquas_main circuit.qasm

Replaces evaluator-preprocess/process.cpp + benchmark/preprocess.sh of the reference (V100 tables, src/evaluator.h:18-117).
Model (ms per 2^30 amplitudes, scaled by 2^(L-30)):
    tile-kernel launch  = max(sweep, group_base + sum(gate cost) + round * extra rounds)            (interpreter kernel, HQ_JIT=0)
                        = max(sweep + 0.015 I, 0.3 + instr * (I + 2 R) + jit_round * (R - 1))          (specialised kernels; I = FP64
                          instructions per amplitude, R = register rounds; see host/evaluator.cpp)
    dense launch        = max(sweep, dense_base + sum(matrix cost[m]))
"""
import json
import sys

GATE_TYPES = ["CCX", "CNOT", "CY", "CZ", "CRX", "CRY", "CU1", "CRZ", "U1", "U2", "U3", "H", "X", "Y", "Z", "S", "SDG", "T",
              "TDG", "RX", "RY", "RZ"]
# microbench case -> gate types priced by it
CASE_TYPES = {"h_x64_4q": ["H"], "rx_x64_4q": ["RX"], "u3_x64_4q": ["U3", "U2"], "t_x64_4q": ["T", "TDG", "S", "SDG", "U1", "Z"],
              "rz_x64_4q": ["RZ"], "x_x64_4q": ["X", "Y"], "cu1fan_x64_4q": ["CZ", "CU1"], "cnot_x64_4q": ["CNOT", "CY", "CCX"]}


def main():
    d = json.load(open(sys.argv[1]))
    scale = 2.0 ** (30 - d["qubits"])
    c = d["cases"]
    sweep = c["sweep_1gate"]["ms_total"] * scale
    h16, h64, h256 = (c[k]["ms_total"] * scale for k in ("h_x16_4q", "h_x64_4q", "h_x256_4q"))
    h_cost = (h256 - h64) / 192.0
    base = h64 - 64 * h_cost
    print(f"hbm_gbs {32.0 * 2**30 / (sweep * 1e-3) / 1e9:.1f}")
    print(f"group_base_ms30 {max(base, 0.0):.3f}")
    for case, types in CASE_TYPES.items():
        if case not in c:
            continue
        cost = max(0.02, (c[case]["ms_total"] * scale - base) / c[case]["gates"])
        for t in types:
            print(f"gate {GATE_TYPES.index(t)} {cost:.4f}")
    # controlled rotations: no dedicated case -> general-mask bodies cost about an uncontrolled rotation
    if "rx_x64_4q" in c:   # (a --only run of the microbenchmark may have skipped it)
        rx = (c["rx_x64_4q"]["ms_total"] * scale - base) / 64
        for t in ("CRX", "CRY", "CRZ", "RY"):
            print(f"gate {GATE_TYPES.index(t)} {rx * (0.75 if t != 'RY' else 0.8):.4f}")
    extra = (c["h_x96_12q"]["ms_total"] * scale - base - 96 * h_cost) / 2.0
    print(f"round_ms30 {max(extra, 0.0):.3f}")
    # specialised (JIT) tile kernels, structural model: ms per FP64 instruction per amplitude (H = 2, U3 = 6 instructions per
    # amplitude) from the long single-round cases; one extra round from the 12-qubit case (3 rounds)
    instr = h256 / (256 * 2.0)
    if "u3_x64_4q" in c:
        instr = max(instr, c["u3_x64_4q"]["ms_total"] * scale / (64 * 6.0))
    print(f"instr_ms30 {instr:.4f}")
    jr = (c["h_x96_12q"]["ms_total"] * scale - 96 * 2.0 * instr) / 2.0
    if jr > 0.1:   # an FP64-bound case hides the exchange cost: keep the evaluator's default (0.9, from supremacy_30 launches) then
        print(f"jit_round_ms30 {jr:.3f}")
    dn = d.get("dense", {})
    if dn:
        m = {k: v["ms"] * scale for k, v in dn.items()}
        dbase = 0.6
        if "m4x2" in m and "m4_high" in m:
            m4 = (m["m4x2"] - dbase) / 2
        else:
            m4 = m.get("m4_high", 6.0) - dbase
        m3 = (m["m3x3"] - dbase) / 3 if "m3x3" in m else m4 / 2
        print(f"dense_base_ms30 {dbase}")
        for q in range(4):
            print(f"dense {q} {m3:.3f}")
        print(f"dense 4 {m4:.3f}")
        print(f"dense 5 {m.get('m5_high', 10.3) - dbase:.3f}")
        print(f"dense 6 {m.get('m6_high', 20.7) - dbase:.3f}")


if __name__ == "__main__":
    main()
