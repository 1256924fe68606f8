// hq_preprocess [L ...]: calibrates the evaluator on THIS GPU and writes, per local-qubit count L, a parameter file in the
// reference's layout (evaluator-preprocess/process.cpp:167-178, read by src/evaluator.cpp:60-103):
//
//   param_type (1 = partial)
//   14 lines: microseconds of a 512-gate circuit of ONE single-qubit type on qubit 1, order U1 U2 U3 H X Y Z S SDG T TDG RX RY RZ
//   7 lines:  same for control 0 -> target 2, order CNOT CY CZ CRX CRY CU1 CRZ
//   10 lines: "K ms" for one dense launch with a K x K matrix, K = 1, 2, ..., 512 (the reference times cublasZgemm there; here
//             the fused dense kernel, K <= 64; larger K, which this build does not run, are written as 8x the previous line)
//   1 line:   transpose cost in ms (cuTT in the reference; the fused kernel has no separate transpose: 0)
//   + a trailing "#hq_preprocess" marker line (after everything the reference's reader consumes)
//
// Output directory: $HYQUAS_PARAM_DIR, else ../evaluator-preprocess/parameter-files (the reference's cwd-relative path), created
// if missing.  Evaluator::loadParam reads these files back (same search order) and recalibrates its model from them.
// Same role as the reference's `process` tool; default L list = its qubit_nums {22..28} (process.cpp:15-16).
#include <sys/stat.h>

#include <complex>
#include <random>
#include <vector>

#include "circuit.h"
#include "logger.h"

static int timeCircuit(Circuit& c) {
    c.compile();
    c.run(false);                 // warm-up (first launch of a freshly built kernel)
    int best = c.run(false);
    for (int i = 0; i < 2; i++) best = std::min(best, c.run(false));
    return best;
}

static double timeDense(int L, int m) {
    const int K = 1 << m;
    std::mt19937_64 rng(7 + m);
    std::normal_distribution<double> nd;
    // a random unitary by Gram-Schmidt (column-major, interleaved re/im)
    std::vector<std::complex<double>> U((size_t)K * K);
    for (auto& v : U) v = {nd(rng), nd(rng)};
    for (int c = 0; c < K; c++) {
        for (int p = 0; p < c; p++) {
            std::complex<double> dot = 0;
            for (int r = 0; r < K; r++) dot += std::conj(U[r + (size_t)p * K]) * U[r + (size_t)c * K];
            for (int r = 0; r < K; r++) U[r + (size_t)c * K] -= dot * U[r + (size_t)p * K];
        }
        double nrm = 0;
        for (int r = 0; r < K; r++) nrm += std::norm(U[r + (size_t)c * K]);
        for (int r = 0; r < K; r++) U[r + (size_t)c * K] /= std::sqrt(nrm);
    }
    std::vector<int> pos(std::max(m, 1));
    for (int i = 0; i < m; i++) pos[i] = i;
    void* st = nullptr;
    checkHq(hq_state_alloc(L, &st));
    checkHq(hq_state_init(st, L, 1));
    hq_dense_plan* plan = nullptr;
    const int mm = std::max(m, 1);   // K = 1: the identity on one qubit
    std::vector<double> flat;
    if (m == 0) flat = {1, 0, 0, 0, 0, 0, 1, 0};
    else for (auto& v : U) { flat.push_back(v.real()); flat.push_back(v.imag()); }
    checkHq(hq_dense_plan_create(L, 1, &mm, pos.data(), flat.data(), &plan));
    float best = 1e30f;
    for (int i = 0; i < 5; i++) {
        checkHq(hq_timer_start());
        checkHq(hq_dense_plan_launch(plan, st, 0));
        float ms = 0;
        checkHq(hq_timer_stop_ms(&ms));
        if (i > 0) best = std::min(best, ms);
    }
    hq_dense_plan_destroy(plan);
    checkHq(hq_state_free(st));
    return best;
}

static void process(int L, const std::string& dir) {
    printf("processing qubit number : %d\n", L);
    const std::string path = dir + "/" + std::to_string(L) + "qubits.out";
    FILE* f = fopen(path.c_str(), "w");
    if (!f) { printf("cannot write %s\n", path.c_str()); exit(1); }
    fprintf(f, "1\n");
    const int numGates = 512;
    for (int i = int(GateType::U1); i < int(GateType::TOTAL); i++) {
        Circuit c(L);
        for (int k = 0; k < numGates; k++) c.addGate(Gate::random(1, 2, GateType(i)));
        fprintf(f, "%d \n", timeCircuit(c));
    }
    fprintf(f, "\n");
    for (int g = int(GateType::CNOT); g <= int(GateType::CRZ); g++) {
        Circuit c(L);
        for (int k = 0; k < numGates; k++) c.addGate(Gate::control(0, 2, GateType(g)));
        fprintf(f, "%d \n", timeCircuit(c));
    }
    double last = 0;
    for (int m = 0; m < 10; m++) {
        if (m <= 6) last = timeDense(L, m); else last *= 8;
        fprintf(f, "%d %f\n", 1 << m, last);
    }
    fprintf(f, "\n%f\n#hq_preprocess (hyquas_b200; the reference's reader stops before this line)\n", 0.0);
    fclose(f);
    Logger::print();
}

int main(int argc, char** argv) {
    setenv("HQ_NUM_GPUS", "1", 1);     // calibration is per GPU
    setenv("HQ_PEEPHOLE", "0", 1);     // time the gates as written (512 identical gates would otherwise be multiplied together)
    setenv("HQ_BACKEND", "group", 1);  // the single / control tables price the tile kernel
    MyGlobalVars::init();
    std::string dir = getenv("HYQUAS_PARAM_DIR") ? getenv("HYQUAS_PARAM_DIR") : "../evaluator-preprocess/parameter-files";
    std::string p;
    for (size_t i = 1; i <= dir.size(); i++)
        if (i == dir.size() || dir[i] == '/') { p = dir.substr(0, i); mkdir(p.c_str(), 0755); }
    std::vector<int> list;
    for (int i = 1; i < argc; i++) list.push_back(atoi(argv[i]));
    if (list.empty()) list = {22, 23, 24, 25, 26, 27, 28};
    for (int L : list) process(L, dir);
    return 0;
}
