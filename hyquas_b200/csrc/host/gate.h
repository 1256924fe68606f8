// Gate IR of the HyQuas-compatible surface: the 22 user gate types + 5 internal ones, each carrying its
// 2x2 matrix.  Same public names and enum order as src/gate.h:7-66 of the reference (the evaluator tool
// iterates GateType by integer value).  The device-side record is hq_gate (include/hyquas_b200.h).
#pragma once

#include <string>
#include <vector>
#include "utils.h"

enum class GateType {
    CCX, CNOT, CY, CZ, CRX, CRY, CU1, CRZ, U1, U2, U3, H, X, Y, Z, S, SDG, T, TDG, RX, RY, RZ, TOTAL, ID, GII, GZZ, GOC, GCC
};

struct Gate {
    int gateID;
    GateType type;
    qComplex mat[2][2];
    std::string name;
    int targetQubit, controlQubit, controlQubit2;   // controls: -1 when absent
    mutable double instrCost = -1.0;                // Evaluator::instrPerAmp's memo (a function of mat and the controls only)
    Gate(): gateID(0), type(GateType::ID), targetQubit(-1), controlQubit(-1), controlQubit2(-1) {}

    bool isControlGate() const { return controlQubit != -1; }
    bool isC2Gate() const { return controlQubit2 != -1; }
    // decided from the matrix (so TDG / ID / GOC ... are covered, whatever their type tag says)
    bool isDiagonal() const { return mat[0][1].x == 0 && mat[0][1].y == 0 && mat[1][0].x == 0 && mat[1][0].y == 0; }

    // factories, by shape: (controls..., target[, angles...]) -- same names and argument order as the reference's
    static Gate CCX(int c1, int c2, int t);
    static Gate CNOT(int c, int t);          static Gate CY(int c, int t);             static Gate CZ(int c, int t);
    static Gate CRX(int c, int t, qreal a);  static Gate CRY(int c, int t, qreal a);   static Gate CRZ(int c, int t, qreal a);
    static Gate CU1(int c, int t, qreal lambda);
    static Gate U1(int t, qreal lambda);     static Gate U2(int t, qreal phi, qreal lambda);
    static Gate U3(int t, qreal theta, qreal phi, qreal lambda);
    static Gate H(int t);    static Gate X(int t);    static Gate Y(int t);    static Gate Z(int t);
    static Gate S(int t);    static Gate SDG(int t);  static Gate T(int t);    static Gate TDG(int t);
    static Gate RX(int t, qreal a);          static Gate RY(int t, qreal a);           static Gate RZ(int t, qreal a);
    // internal gates produced by per-GPU lowering: identity, i*I, -I, diag(1, z), z*I
    static Gate ID(int t);   static Gate GII(int t);  static Gate GZZ(int t);
    static Gate GOC(int t, qreal re, qreal im);       static Gate GCC(int t, qreal re, qreal im);

    static Gate random(int lo, int hi);                 // micro-benchmarks: random type / angles on a qubit in [lo, hi)
    static Gate random(int lo, int hi, GateType type);
    static Gate control(int c, int t, GateType type);   // controlled gate of a given type with random angles
    static GateType toCU(GateType type);                // U -> its controlled form, and back
    static GateType toU(GateType type);
    static std::string get_name(GateType ty);
    // generic constructor used by the factories, the QASM front end and the peephole pass
    static Gate make(GateType type, const char* name, int c2, int c1, int t, const qComplex m[4]);
};
