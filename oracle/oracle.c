/*
 * oracle/oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, gate-by-gate FP64 state-vector replay: the CPU restatement of what the
 * reference computes for a gate list, one full pass over the 2^n amplitudes per gate,
 * no fusion, no blocking.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product path
 * (hyquas_b200/) never does.
 *
 * Reference semantics restated here (paths relative to /root/reference):
 *   - pair index rule      src/kernelSimple.cu:40-70   lo = ((i>>t)<<(t+1)) | (i & mask), hi = lo | 1<<t,
 *                                                      control test on bits of lo
 *   - 2x2 update           src/kernelSimple.cu:187-198 lo' = m00*lo + m01*hi ; hi' = m10*lo + m11*hi
 *   - qubit k <-> bit k of the amplitude index (LSB = q[0])   src/kernelSimple.cu:42-45
 *   - initial state |0..0>  src/kernelSimple.cu:27-30
 *   - timing covers the gate loop only (mirrors "Time Cost")  src/circuit.cpp:22,46,51-52
 *
 * Parity pinning: checked against the reference's own golden outputs qft_28.log, bv_28.log,
 * hidden_shift_28.log (tests/golden/, sha256 == the Git-LFS oids of /root/reference/tests/output).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } amp_t;

/* One gate record as passed from Python: matrix is row-major (m00,m01,m10,m11) x (re,im). */
typedef struct {
    int32_t target;
    int32_t control;   /* -1 if none */
    int32_t control2;  /* -1 if none */
    int32_t pad;
    double  m[8];
} orc_gate;

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_init_zero_state(amp_t* s, int n) {
    int64_t N = (int64_t)1 << n;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; i++) { s[i].re = 0.0; s[i].im = 0.0; }
    s[0].re = 1.0;
}

static inline int is_diag(const double* m) {
    return m[2] == 0.0 && m[3] == 0.0 && m[4] == 0.0 && m[5] == 0.0;
}

/* General (possibly controlled) 2x2 update, kernelSimple.cu:40-70 + :187-198. */
void orc_apply_gate(amp_t* s, int n, const orc_gate* g) {
    const int t = g->target, c = g->control, c2 = g->control2;
    const int64_t half = (int64_t)1 << (n - 1);
    const int64_t mask = ((int64_t)1 << t) - 1;
    const int64_t tbit = (int64_t)1 << t;
    const double r00 = g->m[0], i00 = g->m[1], r01 = g->m[2], i01 = g->m[3];
    const double r10 = g->m[4], i10 = g->m[5], r11 = g->m[6], i11 = g->m[7];
    const int diag = is_diag(g->m);
    const int lo_unit = (r00 == 1.0 && i00 == 0.0);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < half; i++) {
        int64_t lo = ((i >> t) << (t + 1)) | (i & mask);
        if (c >= 0 && !((lo >> c) & 1)) continue;
        if (c2 >= 0 && !((lo >> c2) & 1)) continue;
        int64_t hi = lo | tbit;
        if (diag) {
            /* specialised diagonal update (allowed by BASELINE.md 3b): only the non-unit entries are touched */
            if (!lo_unit) {
                double x = s[lo].re, y = s[lo].im;
                s[lo].re = x * r00 - y * i00;
                s[lo].im = x * i00 + y * r00;
            }
            double x = s[hi].re, y = s[hi].im;
            s[hi].re = x * r11 - y * i11;
            s[hi].im = x * i11 + y * r11;
        } else {
            double lr = s[lo].re, li = s[lo].im, hr = s[hi].re, hi_ = s[hi].im;
            s[lo].re = (lr * r00 - li * i00) + (hr * r01 - hi_ * i01);
            s[lo].im = (lr * i00 + li * r00) + (hr * i01 + hi_ * r01);
            s[hi].re = (lr * r10 - li * i10) + (hr * r11 - hi_ * i11);
            s[hi].im = (lr * i10 + li * r10) + (hr * i11 + hi_ * r11);
        }
    }
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* Replay a gate list; returns seconds spent in the gate loop only. */
double orc_run(amp_t* s, int n, const orc_gate* gates, int ngates) {
    double t0 = now_s();
    for (int k = 0; k < ngates; k++) orc_apply_gate(s, n, &gates[k]);
    return now_s() - t0;
}

/* Helpers for dumps without touching all of memory from Python. */
int64_t orc_scan_large(const amp_t* s, int n, double thresh, int64_t min_idx, int64_t* out_idx, int64_t cap) {
    int64_t N = (int64_t)1 << n, cnt = 0;
    for (int64_t i = min_idx; i < N; i++) {
        double p = s[i].re * s[i].re + s[i].im * s[i].im;
        if (p > thresh) { if (cnt < cap) out_idx[cnt] = i; cnt++; }
    }
    return cnt;
}

double orc_norm2(const amp_t* s, int n) {
    int64_t N = (int64_t)1 << n;
    double acc = 0.0;
#pragma omp parallel for reduction(+:acc) schedule(static)
    for (int64_t i = 0; i < N; i++) acc += s[i].re * s[i].re + s[i].im * s[i].im;
    return acc;
}
