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
    int targetQubit;
    int controlQubit;   // -1 if no control
    int controlQubit2;  // -1 if no second control
    Gate(): gateID(0), type(GateType::ID), targetQubit(-1), controlQubit(-1), controlQubit2(-1) {}
    bool isControlGate() const { return controlQubit != -1; }
    bool isC2Gate() const { return controlQubit2 != -1; }
    // true when the matrix is diagonal (decided from the matrix, so TDG/ID/GOC... are covered too)
    bool isDiagonal() const {
        return mat[0][1].x == 0 && mat[0][1].y == 0 && mat[1][0].x == 0 && mat[1][0].y == 0;
    }
    static Gate CCX(int c1, int c2, int targetQubit);
    static Gate CNOT(int controlQubit, int targetQubit);
    static Gate CY(int controlQubit, int targetQubit);
    static Gate CZ(int controlQubit, int targetQubit);
    static Gate CRX(int controlQubit, int targetQubit, qreal angle);
    static Gate CRY(int controlQubit, int targetQubit, qreal angle);
    static Gate CU1(int controlQubit, int targetQubit, qreal lambda);
    static Gate CRZ(int controlQubit, int targetQubit, qreal angle);
    static Gate U1(int targetQubit, qreal lambda);
    static Gate U2(int targetQubit, qreal phi, qreal lambda);
    static Gate U3(int targetQubit, qreal theta, qreal phi, qreal lambda);
    static Gate H(int targetQubit);
    static Gate X(int targetQubit);
    static Gate Y(int targetQubit);
    static Gate Z(int targetQubit);
    static Gate S(int targetQubit);
    static Gate SDG(int targetQubit);
    static Gate T(int targetQubit);
    static Gate TDG(int targetQubit);
    static Gate RX(int targetQubit, qreal angle);
    static Gate RY(int targetQubit, qreal angle);
    static Gate RZ(int targetQubit, qreal angle);
    static Gate ID(int targetQubit);
    static Gate GII(int targetQubit);
    static Gate GZZ(int targetQubit);
    static Gate GOC(int targetQubit, qreal real, qreal imag);
    static Gate GCC(int targetQubit, qreal real, qreal imag);
    static Gate random(int lo, int hi);
    static Gate random(int lo, int hi, GateType type);
    static Gate control(int controlQubit, int targetQubit, GateType type);
    static GateType toCU(GateType type);
    static GateType toU(GateType type);
    static std::string get_name(GateType ty);
    // generic constructor used by the factories and by the QASM front end
    static Gate make(GateType type, const char* name, int c2, int c1, int t, const qComplex m[4]);
};
