// Host-side basics of the HyQuas-compatible C++ surface: scalar/index/complex typedefs, tunables, the
// per-process globals (MyGlobalVars, MyMPI) and the abort-on-error convention.
// Mirrors the names a HyQuas user sees in src/utils.h:16-61,146-164; everything CUDA-specific lives
// behind the C-ABI in include/hyquas_b200.h (no streams / cuBLAS / cuTT handles here).
#pragma once

#include <cstdio>
#include <cstdlib>
#include <cuComplex.h>
#include <memory>
#include <string>

#include "hyquas_b200.h"

typedef double qreal;                 // FP64 only (the reference's USE_DOUBLE=on build)
typedef long long qindex;
typedef cuDoubleComplex qComplex;     // {double x, y}
#define make_qComplex make_cuDoubleComplex

#ifndef BLAS_MAT_LIMIT_DEFINED
#define BLAS_MAT_LIMIT_DEFINED 6
#endif
const int LOCAL_QUBIT_SIZE = 10;      // kept for source compatibility (the micro-benchmark drivers loop over it);
                                      // the real tile width is hq_group_tile_bits()
const int BLAS_MAT_LIMIT = BLAS_MAT_LIMIT_DEFINED;
const int MIN_MAT_SIZE = 4;
const int COALESCE_GLOBAL = 3;
const int MAX_GATE = 600;

#define UNREACHABLE() { \
    printf("file %s line %i: unreachable!\n", __FILE__, __LINE__); \
    fflush(stdout); \
    exit(1); \
}

// Abort-on-error wrapper for the C-ABI (reference convention: print and exit(1), src/utils.h:106-144).
#define checkHq(stmt) { \
    int hq_rc_ = (stmt); \
    if (hq_rc_ != HQ_OK) { \
        fprintf(stderr, "%s in file %s, function %s, line %i: %04d %s\n", #stmt, __FILE__, __FUNCTION__, __LINE__, hq_rc_, hq_last_error()); \
        exit(1); \
    } \
}

namespace MyGlobalVars {
    extern int numGPUs;    // GPUs taking part in the simulation (= number of processes, one GPU each)
    extern int localGPUs;  // GPUs driven by this process: always 1
    extern int bit;        // log2(numGPUs) = number of global qubits
    extern bool swapAnyBit; // swaps may trade ANY local position >= 3 (p2p transport); else only the top k positions
    extern bool hostOnly;  // set by initForTest(): no GPU bound, plans stay on the host
    void init();
    void initForTest(int worldSize, int rank);   // host-only: no GPU is touched (compiler / plan tests)
}

namespace MyMPI {
    extern int rank;       // this process' GPU index in [0, numGPUs)
    extern int commSize;
    extern int commBit;
    void init();
}

template<typename T>
int bitCount(T x) {
    int ret = 0;
    for (; x; x &= x - 1) ret++;
    return ret;
}

qreal zero_wrapper(qreal x);
qComplex operator * (const qComplex& a, const qComplex& b);
qComplex operator + (const qComplex& a, const qComplex& b);
qComplex make_qComplex(qreal x);
bool operator < (const qComplex& a, const qComplex& b);
int get_bit(int n);
