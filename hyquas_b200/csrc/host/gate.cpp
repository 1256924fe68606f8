// Gate factories.  Matrix entries follow the reference's definitions (src/gate.cpp:9-342) expression by
// expression so that the lowered 2x2 blocks are bit-identical to what the reference feeds its kernels:
//   X/CNOT/CCX [[0,1],[1,0]] (:9-33,164-174)      Y/CY [[0,-i],[i,0]] (:35-45,176-186)   Z/CZ diag(1,-1) (:47-57,188-198)
//   RX/CRX [[c,-is],[-is,c]], c=cos(a/2) (:59-69,248-258)     RY/CRY [[c,-s],[s,c]] (:71-81,260-270)
//   RZ/CRZ diag(e^{-ia/2}, e^{ia/2}) (:97-107,272-282)        U1/CU1 diag(1, e^{il}) (:83-95,110-122)
//   U2 (:124-136)   U3 (:138-150)   H (:152-162)   S/SDG (:200-222)   T/TDG (:224-246)
//   ID, GII = i*I, GZZ = -I, GOC = diag(1,z), GCC = z*I (:284-342)
// One deliberate difference: TDG is tagged GateType::TDG (the reference tags it T, gate.cpp:239, which
// makes its shared-memory kernel apply T for tdg); the matrix is the same correct diag(1, e^{-i pi/4}).
#include "gate.h"

#include <cassert>
#include <cmath>

static int nextGateID = 0;

Gate Gate::make(GateType type, const char* name, int c2, int c1, int t, const qComplex m[4]) {
    Gate g;
    g.gateID = ++nextGateID;
    g.type = type;
    g.name = name;
    g.controlQubit2 = c2;
    g.controlQubit = c1;
    g.targetQubit = t;
    g.mat[0][0] = m[0]; g.mat[0][1] = m[1]; g.mat[1][0] = m[2]; g.mat[1][1] = m[3];
    return g;
}

namespace {
typedef qComplex C;
inline C c(qreal re, qreal im = 0.0) { return make_cuDoubleComplex(re, im); }

struct M4 { C m[4]; };
M4 mX() { return {{c(0), c(1), c(1), c(0)}}; }
M4 mY() { return {{c(0), c(0, -1), c(0, 1), c(0)}}; }
M4 mZ() { return {{c(1), c(0), c(0), c(-1)}}; }
M4 mRX(qreal a) { return {{c(cos(a / 2.0)), c(0, -sin(a / 2.0)), c(0, -sin(a / 2.0)), c(cos(a / 2.0))}}; }
M4 mRY(qreal a) { return {{c(cos(a / 2.0)), c(-sin(a / 2.0)), c(sin(a / 2.0)), c(cos(a / 2.0))}}; }
M4 mRZ(qreal a) { return {{c(cos(a / 2), -sin(a / 2)), c(0), c(0), c(cos(a / 2), sin(a / 2))}}; }
M4 mU1(qreal l) { return {{c(1), c(0), c(0), c(cos(l), sin(l))}}; }
M4 mDiag(C d0, C d1) { return {{d0, c(0), c(0), d1}}; }

Gate build(GateType ty, const char* name, const M4& m, int t, int c1 = -1, int c2 = -1) {
    return Gate::make(ty, name, c2, c1, t, m.m);
}
}  // namespace

Gate Gate::CCX(int c1, int c2, int t) { return build(GateType::CCX, "CCX", mX(), t, c1, c2); }
Gate Gate::CNOT(int cq, int t) { return build(GateType::CNOT, "CN", mX(), t, cq); }
Gate Gate::CY(int cq, int t) { return build(GateType::CY, "CY", mY(), t, cq); }
Gate Gate::CZ(int cq, int t) { return build(GateType::CZ, "CZ", mZ(), t, cq); }
Gate Gate::CRX(int cq, int t, qreal a) { return build(GateType::CRX, "CRX", mRX(a), t, cq); }
Gate Gate::CRY(int cq, int t, qreal a) { return build(GateType::CRY, "CRY", mRY(a), t, cq); }
Gate Gate::CU1(int cq, int t, qreal l) { return build(GateType::CU1, "CU1", mU1(l), t, cq); }
Gate Gate::CRZ(int cq, int t, qreal a) { return build(GateType::CRZ, "CRZ", mRZ(a), t, cq); }
Gate Gate::U1(int t, qreal l) { return build(GateType::U1, "U1", mU1(l), t); }
Gate Gate::U2(int t, qreal phi, qreal lambda) {
    M4 m = {{c(1.0 / sqrt(2)), c(-cos(lambda) / sqrt(2), -sin(lambda) / sqrt(2)),
             c(cos(phi) / sqrt(2), sin(phi) / sqrt(2)), c(cos(lambda + phi) / sqrt(2), sin(lambda + phi) / sqrt(2))}};
    return build(GateType::U2, "U2", m, t);
}
Gate Gate::U3(int t, qreal theta, qreal phi, qreal lambda) {
    M4 m = {{c(cos(theta / 2)), c(-cos(lambda) * sin(theta / 2), -sin(lambda) * sin(theta / 2)),
             c(cos(phi) * sin(theta / 2), sin(phi) * sin(theta / 2)),
             c(cos(phi + lambda) * cos(theta / 2), sin(phi + lambda) * cos(theta / 2))}};
    return build(GateType::U3, "U3", m, t);
}
Gate Gate::H(int t) {
    M4 m = {{c(1 / sqrt(2)), c(1 / sqrt(2)), c(1 / sqrt(2)), c(-1 / sqrt(2))}};
    return build(GateType::H, "H", m, t);
}
Gate Gate::X(int t) { return build(GateType::X, "X", mX(), t); }
Gate Gate::Y(int t) { return build(GateType::Y, "Y", mY(), t); }
Gate Gate::Z(int t) { return build(GateType::Z, "Z", mZ(), t); }
Gate Gate::S(int t) { return build(GateType::S, "S", mDiag(c(1), c(0, 1)), t); }
Gate Gate::SDG(int t) { return build(GateType::SDG, "SDG", mDiag(c(1), c(0, -1)), t); }
Gate Gate::T(int t) { return build(GateType::T, "T", mDiag(c(1), c(1 / sqrt(2), 1 / sqrt(2))), t); }
Gate Gate::TDG(int t) { return build(GateType::TDG, "TDG", mDiag(c(1), c(1 / sqrt(2), -1 / sqrt(2))), t); }
Gate Gate::RX(int t, qreal a) { return build(GateType::RX, "RX", mRX(a), t); }
Gate Gate::RY(int t, qreal a) { return build(GateType::RY, "RY", mRY(a), t); }
Gate Gate::RZ(int t, qreal a) { return build(GateType::RZ, "RZ", mRZ(a), t); }
Gate Gate::ID(int t) { return build(GateType::ID, "ID", mDiag(c(1), c(1)), t); }
Gate Gate::GII(int t) { return build(GateType::GII, "GII", mDiag(c(0, 1), c(0, 1)), t); }
Gate Gate::GZZ(int t) { return build(GateType::GZZ, "GZZ", mDiag(c(-1), c(-1)), t); }
Gate Gate::GOC(int t, qreal re, qreal im) { return build(GateType::GOC, "GOC", mDiag(c(1), c(re, im)), t); }
Gate Gate::GCC(int t, qreal re, qreal im) { return build(GateType::GCC, "GCC", mDiag(c(re, im), c(re, im)), t); }

// ---- random instances (micro-benchmarks / evaluator calibration, reference gate.cpp:344-521) --------
namespace {
qreal rand01() { return rand() * 1.0 / RAND_MAX; }
qreal randAngle() { return rand01() * acos(-1) * 2; }
int pick(int lo, int hi) { return rand() % (hi - lo) + lo; }
}  // namespace

Gate Gate::random(int lo, int hi) { return random(lo, hi, GateType(rand() % int(GateType::TOTAL))); }

Gate Gate::random(int lo, int hi, GateType type) {
    const int arity = type == GateType::CCX ? 3 : (int(type) <= int(GateType::CRZ) ? 2 : 1);
    assert(hi - lo >= arity);
    int q[3] = {-1, -1, -1};   // controls first, target last drawn (c2, c1, t order as the reference draws them)
    for (;;) {
        for (int i = 0; i < arity; i++) q[i] = pick(lo, hi);
        bool distinct = true;
        for (int i = 0; i < arity; i++) for (int j = 0; j < i; j++) distinct &= q[i] != q[j];
        if (distinct) break;
    }
    if (arity == 3) return CCX(q[1], q[0], q[2]);
    if (arity == 2) {
        const int cq = q[0], t = q[1];
        switch (type) {
            case GateType::CNOT: return CNOT(cq, t);
            case GateType::CY: return CY(cq, t);
            case GateType::CZ: return CZ(cq, t);
            case GateType::CRX: return CRX(cq, t, randAngle());
            case GateType::CRY: return CRY(cq, t, randAngle());
            case GateType::CU1: return CU1(cq, t, randAngle());
            default: return CRZ(cq, t, randAngle());
        }
    }
    const int t = q[0];
    switch (type) {
        case GateType::U1: return U1(t, randAngle());
        case GateType::U2: { qreal a = randAngle(), b = randAngle(); return U2(t, a, b); }
        case GateType::U3: { qreal a = randAngle(), b = randAngle(), d = randAngle(); return U3(t, a, b, d); }
        case GateType::H: return H(t);
        case GateType::X: return X(t);
        case GateType::Y: return Y(t);
        case GateType::Z: return Z(t);
        case GateType::S: return S(t);
        case GateType::SDG: return SDG(t);
        case GateType::T: return T(t);
        case GateType::TDG: return TDG(t);
        case GateType::RX: return RX(t, randAngle());
        case GateType::RY: return RY(t, randAngle());
        case GateType::RZ: return RZ(t, randAngle());
        default:
            printf("invalid %d\n", (int)type);
            exit(1);
    }
}

Gate Gate::control(int cq, int t, GateType type) {
    switch (type) {
        case GateType::CNOT: return CNOT(cq, t);
        case GateType::CY: return CY(cq, t);
        case GateType::CZ: return CZ(cq, t);
        case GateType::CRX: return CRX(cq, t, randAngle());
        case GateType::CRY: return CRY(cq, t, randAngle());
        case GateType::CU1: return CU1(cq, t, randAngle());
        case GateType::CRZ: return CRZ(cq, t, randAngle());
        default: UNREACHABLE()
    }
}

GateType Gate::toCU(GateType type) {
    if (type == GateType::CCX) return GateType::CNOT;
    UNREACHABLE()
}

GateType Gate::toU(GateType type) {
    switch (type) {
        case GateType::CCX: case GateType::CNOT: return GateType::X;
        case GateType::CY: return GateType::Y;
        case GateType::CZ: return GateType::Z;
        case GateType::CRX: return GateType::RX;
        case GateType::CRY: return GateType::RY;
        case GateType::CU1: return GateType::U1;
        case GateType::CRZ: return GateType::RZ;
        default: UNREACHABLE()
    }
}

std::string Gate::get_name(GateType ty) {
    static const char* names[] = {"CCX", "CN", "CY", "CZ", "CRX", "CRY", "CU1", "CRZ", "U1", "U2", "U3", "H", "X", "Y", "Z",
                                  "S", "SDG", "T", "TDG", "RX", "RY", "RZ", "TOTAL", "ID", "GII", "GZZ", "GOC", "GCC"};
    return names[int(ty)];
}
