#include "evaluator.h"

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

Evaluator* Evaluator::getInstance() {
    static Evaluator* inst = new Evaluator();
    return inst;
}

Evaluator::Evaluator() {
    // Defaults = this pool's B200, FP64, measured with tools/microbench.py at 2^30 amplitudes (profiles/r01_s16_microbench.json):
    // a gate-group launch of G same-type gates takes  max(sweep, groupBaseMs30 + G * gate cost + extra rounds * roundMs30);
    // a fused dense launch takes  max(sweep, denseBaseMs30 + sum of per-matrix costs).  tools/calibrate.py rewrites them
    // from a fresh microbenchmark run ($HYQUAS_PARAM_FILE).
    hbmGBs = 6250.0;        // one in-place sweep (16 B read + 16 B written per amplitude): 5.5 ms per 2^30
    launchMs = 0.01;
    nvlinkGBs = 700.0;
    groupBaseMs30 = 2.6;
    denseBaseMs30 = 0.6;
    circuitFactor = 1.45;
    roundMs30 = 2.0;        // h_x96_12q gives 1.2; groups of real circuits (predicates on lane bits, uneven rounds) fit 2.0
    for (auto& g : gateNs) g = 0.35;
    auto set = [&](GateType t, double ms30) { gateNs[int(t)] = ms30; };
    // H prices every butterfly (H, RX/RY(+-pi/2): two FP64 adds per amplitude); see perfPerGate(gates)
    set(GateType::H, 0.18); set(GateType::RY, 0.35); set(GateType::RX, 0.43);
    set(GateType::U2, 0.68); set(GateType::U3, 0.68);
    // diag(1,d) gates merge into per-register-bit diagonal runs (one factor per thread)
    set(GateType::T, 0.09); set(GateType::TDG, 0.09); set(GateType::S, 0.09); set(GateType::SDG, 0.09); set(GateType::U1, 0.09);
    // (RZ = scalar * diag(1, e^{ia}): the tile kernel defers the scalar, so it prices like U1)
    set(GateType::RZ, 0.09); set(GateType::Z, 0.09);
    set(GateType::GOC, 0.09); set(GateType::ID, 0.0); set(GateType::GII, 0.0); set(GateType::GZZ, 0.0); set(GateType::GCC, 0.0); set(GateType::X, 0.43); set(GateType::Y, 0.43);
    set(GateType::CZ, 0.10); set(GateType::CU1, 0.10); set(GateType::CRZ, 0.25);
    set(GateType::CNOT, 0.25); set(GateType::CY, 0.27); set(GateType::CCX, 0.25);
    set(GateType::CRX, 0.33); set(GateType::CRY, 0.33);
    // specialised kernels (profiles/r02_s1_microbench_jit.json): 64..256 butterflies run at 0.134-0.138 ms each (2 FP64
    // instructions per amplitude, 85-92 % of the 18.6 T instr/s peak), 64 U3 at 0.356 ms each (6 per amplitude); permutations,
    // CZ / Z / S and diagonal runs on thread bits ride the 5.5 ms sweep
    int avail = 0;
    hq_jit_available(&avail);
    specialised = avail != 0;
    instrMs30 = 0.060;      // h_x256_4q: 512 instructions per amplitude in 30.8 ms
    jitRoundMs30 = 0.9;     // supremacy_30 launches: 112-124 instructions per amplitude in 3-5 rounds take 9.6-11.6 ms
    jitBaseMs30 = 0.3;
    jitUnderSweepMs30 = 0.015;
    // r02_s8, quantum_volume_30: 384.8 ms priced gate by gate (40 dense launches), 375.9 ms fusion-aware (28 dense + 11 tile)
    fusionAware = !(getenv("HQ_EVAL_FUSION") != nullptr && atoi(getenv("HQ_EVAL_FUSION")) == 0);
    const double dense[8] = {2.75, 2.75, 2.75, 2.75, 5.0, 9.7, 18.3, 41.0};   // by matrix qubits (<= 3 padded to 3; 7 not built)
    for (int m = 0; m < 8; m++) denseMs30[m] = dense[m];
}

// Parameter file in the reference's layout (evaluator-preprocess/process.cpp:167-178, src/evaluator.cpp:60-103), written for
// this GPU by hq_preprocess: param_type, 14 single-gate + 7 controlled-gate times (us per 512 gates), 10 "K ms" dense lines,
// one transpose line.  Searched in $HYQUAS_PARAM_DIR, then ../evaluator-preprocess/parameter-files (the reference's path).
bool Evaluator::loadReferenceLayout(int numQubits) {
    std::vector<std::string> dirs;
    if (const char* d = getenv("HYQUAS_PARAM_DIR")) dirs.push_back(d);
    dirs.push_back("../evaluator-preprocess/parameter-files");
    for (const std::string& d : dirs) {
        std::ifstream in(d + "/" + std::to_string(numQubits) + "qubits.out");
        if (!in) continue;
        int type = -1;
        double single[14], ctr[7], dense[10], tr = 0;
        in >> type;
        if (type != 1) continue;   // only the partial layout is written / read here
        for (double& v : single) in >> v;
        for (double& v : ctr) in >> v;
        for (double& v : dense) { int K; in >> K >> v; }
        in >> tr;
        if (!in) continue;
        // the default path may hold the REFERENCE's files (its own kernels, possibly another GPU): only files that carry
        // hq_preprocess's trailing marker are taken from there; an explicit $HYQUAS_PARAM_DIR is trusted as it is
        std::string marker;
        in >> marker;
        if (&d != &dirs[0] || !getenv("HYQUAS_PARAM_DIR")) if (marker != "#hq_preprocess") continue;
        const double to30 = std::ldexp(1.0, 30 - numQubits) / 1000.0 / 512.0;   // us per 512 gates at 2^L -> ms per gate at 2^30
        const GateType singles[14] = {GateType::U1, GateType::U2, GateType::U3, GateType::H, GateType::X, GateType::Y, GateType::Z,
                                      GateType::S, GateType::SDG, GateType::T, GateType::TDG, GateType::RX, GateType::RY, GateType::RZ};
        const GateType ctrs[7] = {GateType::CNOT, GateType::CY, GateType::CZ, GateType::CRX, GateType::CRY, GateType::CU1, GateType::CRZ};
        for (int i = 0; i < 14; i++) gateNs[int(singles[i])] = single[i] * to30;
        for (int i = 0; i < 7; i++) gateNs[int(ctrs[i])] = ctr[i] * to30;
        gateNs[int(GateType::CCX)] = gateNs[int(GateType::CNOT)];
        // structural model of the specialised kernels: H = 2 FP64 instructions per amplitude, U3 = 6
        instrMs30 = std::max(single[3] * to30 / 2.0, single[2] * to30 / 6.0);
        for (int m = 3; m <= 6; m++) denseMs30[m] = dense[m] * std::ldexp(1.0, 30 - numQubits);
        for (int m = 0; m < 3; m++) denseMs30[m] = denseMs30[3];
        return true;
    }
    return false;
}

void Evaluator::loadParam(int numQubits) {
    if (!triedLayout.count(numQubits)) {
        triedLayout.insert(numQubits);
        loadReferenceLayout(numQubits);
    }
    if (loaded) return;
    loaded = true;
    const char* path = getenv("HYQUAS_PARAM_FILE");
    if (!path) return;
    std::ifstream in(path);
    if (!in) {
        printf("Parameter file not find: %s\n", path);
        exit(1);
    }
    // "key value" lines: hbm_gbs, launch_ms, round_ms30, gate <GateType index> <ms30>, dense <m> <ms30>
    std::string key;
    while (in >> key) {
        if (key == "hbm_gbs") in >> hbmGBs;
        else if (key == "launch_ms") in >> launchMs;
        else if (key == "nvlink_gbs") in >> nvlinkGBs;
        else if (key == "round_ms30") in >> roundMs30;
        else if (key == "circuit_factor") in >> circuitFactor;
        else if (key == "group_base_ms30") in >> groupBaseMs30;
        else if (key == "instr_ms30") in >> instrMs30;
        else if (key == "jit_round_ms30") in >> jitRoundMs30;
        else if (key == "jit_base_ms30") in >> jitBaseMs30;
        else if (key == "jit_under_sweep_ms30") in >> jitUnderSweepMs30;
        else if (key == "specialised") { int v; in >> v; specialised = v != 0; }
        else if (key == "dense_base_ms30") in >> denseBaseMs30;
        else if (key == "gate") { int i; double v; in >> i >> v; if (i >= 0 && i < 32) gateNs[i] = v; }
        else if (key == "dense") { int i; double v; in >> i >> v; if (i >= 0 && i < 8) denseMs30[i] = v; }
    }
}

double Evaluator::perfPerGate(int numQubits, const std::vector<GateType>& types) {
    loadParam(numQubits);
    double compute = groupBaseMs30;
    for (GateType t : types) compute += gateNs[int(t) & 31];
    return launchMs + std::ldexp(1.0, numQubits - 30) * std::max(sweepMs30(), compute);
}

// alpha * [[1, p], [q, -p q]] with p, q both in {+-1} or both in {+-i}, uncontrolled: the tile kernel's butterfly class
// (same test as classify() in device/group_kernel.cu): H, RX(+-pi/2), RY(+-pi/2) whatever their type tag says.
static bool isButterflyMat(const Gate& g);
static bool isButterfly(const Gate& g) {
    if (g.controlQubit != -1 || g.controlQubit2 != -1) return false;
    return isButterflyMat(g);
}
static bool isButterflyMat(const Gate& g) {
    typedef std::complex<double> C;
    const C a(g.mat[0][0].x, g.mat[0][0].y), b(g.mat[0][1].x, g.mat[0][1].y), c(g.mat[1][0].x, g.mat[1][0].y), d(g.mat[1][1].x, g.mat[1][1].y);
    if (std::abs(a) < 0.5) return false;
    const C p = b / a, q = c / a, r = d / a;
    auto unit = [](C z, bool imag) { return std::abs(std::abs(imag ? z.imag() : z.real()) - 1.0) < 1e-14 && std::abs(imag ? z.real() : z.imag()) < 1e-14; };
    return ((unit(p, false) && unit(q, false)) || (unit(p, true) && unit(q, true))) && std::abs(r + p * q) < 1e-14;
}

// FP64 instructions per amplitude the specialised kernel emits for this gate (device/group_jit.cpp: every scalar is
// coefficient x variable, a sum of n terms costs n - 1 instructions, products by constants are free until the round's flush).
double Evaluator::instrPerAmp(const Gate& g) {
    if (g.instrCost >= 0) return g.instrCost;
    return g.instrCost = instrPerAmpUncached(g);
}

double Evaluator::instrPerAmpUncached(const Gate& g) {
    int nz[2] = {0, 0};
    bool unit = true;   // every non-zero entry is +-1 or +-i
    for (int r = 0; r < 2; r++)
        for (int c = 0; c < 2; c++) {
            const qComplex z = g.mat[r][c];
            if (z.x == 0.0 && z.y == 0.0) continue;
            nz[r]++;
            if (!((std::fabs(z.x) == 1.0 && z.y == 0.0) || (z.x == 0.0 && std::fabs(z.y) == 1.0))) unit = false;
        }
    double cost;
    if (nz[0] <= 1 && nz[1] <= 1) {
        // permutation / diagonal: a non-unit entry (x + iy)(a + ib) is one instruction per scalar of the half it multiplies
        // (plus its share of the round's flush)
        if (unit) cost = 0.02;
        else {
            const bool d0one = g.mat[0][0].x == 1.0 && g.mat[0][0].y == 0.0 && g.mat[0][1].x == 0.0 && g.mat[0][1].y == 0.0;
            cost = d0one ? 1.0 : 2.0;
        }
    } else {
        // dense 2x2: real, RX-like (and butterflies): each output scalar is a sum of 2 terms; general complex: 4 terms
        bool realLike = true, rxLike = true;
        for (int r = 0; r < 2; r++)
            for (int c = 0; c < 2; c++) {
                const qComplex z = g.mat[r][c];
                if (z.y != 0.0) realLike = false;
                if ((r == c ? z.y : z.x) != 0.0) rxLike = false;
            }
        cost = (realLike || rxLike || isButterflyMat(g)) ? 2.0 : 6.0;
    }
    if (g.controlQubit >= 0) cost *= 0.5;
    if (g.controlQubit2 >= 0) cost *= 0.5;
    return cost;
}

// What block fusion in the specialised kernels (device/group_jit.cpp: consecutive static ops within one or two register qubits
// are emitted as their 2x2 / 4x4 product when that is shorter) can bring the instruction count down to, on logical qubits and
// ignoring rounds, i.e. optimistically: a one-qubit block costs at most 6 instructions per amplitude, a two-qubit block 14.
double Evaluator::fusedInstr(const std::vector<Gate>& gates) {
    struct Block { qindex bits; double eager; };
    std::vector<Block> open;
    double total = 0;
    auto close = [&](size_t i) {
        const int nb = bitCount(open[i].bits);
        total += std::min(open[i].eager, nb <= 1 ? 6.0 : 14.0);
        open.erase(open.begin() + i);
    };
    for (const Gate& g : gates) {
        qindex S = qindex(1) << g.targetQubit;
        if (g.controlQubit >= 0) S |= qindex(1) << g.controlQubit;
        if (g.controlQubit2 >= 0) S |= qindex(1) << g.controlQubit2;
        const double c = instrPerAmp(g);
        qindex uni = S;
        for (const Block& b : open) if (b.bits & S) uni |= b.bits;
        if (bitCount(uni) > 2) {
            for (size_t i = 0; i < open.size();) { if (open[i].bits & S) close(i); else ++i; }
            if (bitCount(S) > 2) { total += c; continue; }
            uni = S;
        }
        Block nb{uni, c};
        for (size_t i = 0; i < open.size();) {
            if (open[i].bits & S) { nb.eager += open[i].eager; open.erase(open.begin() + i); }
            else ++i;
        }
        open.push_back(nb);
    }
    while (!open.empty()) close(0);
    return total;
}

// Register rounds the tile kernel's planner will need (device/group_kernel.cu: a round holds 4 register qubits; a gate joins the
// current round when nothing it fails to commute with was left behind and its non-diagonal target is, or can still become, a
// register qubit): the same greedy fill, on logical qubits.
int Evaluator::registerRounds(const std::vector<Gate>& gates) {
    std::vector<char> done(gates.size(), 0);
    size_t left = gates.size();
    int rounds = 0;
    while (left > 0) {
        qindex blockedX = 0, blockedZ = 0, reg = 0;
        int nreg = 0;
        for (size_t i = 0; i < gates.size(); i++) {
            if (done[i]) continue;
            const Gate& g = gates[i];
            const bool diag = g.isDiagonal();
            qindex qn = 0, qd = 0;
            (diag ? qd : qn) |= qindex(1) << g.targetQubit;
            if (g.controlQubit >= 0) qd |= qindex(1) << g.controlQubit;
            if (g.controlQubit2 >= 0) qd |= qindex(1) << g.controlQubit2;
            bool can = !(qn & (blockedX | blockedZ)) && !(qd & blockedX);
            if (can && !diag && !(reg >> g.targetQubit & 1)) {
                if (nreg < 4) { reg |= qindex(1) << g.targetQubit; nreg++; }
                else can = false;
            }
            if (can) { done[i] = 1; left--; }
            else { blockedX |= qn; blockedZ |= qd; }
        }
        rounds++;
    }
    return std::max(1, rounds);
}

double Evaluator::perfPerGate(int numQubits, const std::vector<Gate>& gates) {
    loadParam(numQubits);
    if (specialised) {
        double instr = 0;
        for (const Gate& g : gates) instr += instrPerAmp(g);
        if (fusionAware) instr = 0.25 * instr + 0.75 * fusedInstr(gates);
        const int rounds = registerRounds(gates);
        // every round after the first moves the tile through shared memory once more (64 KB out, 64 KB in per tile: 0.9 ms per
        // 2^30 amplitudes at 128 B/clk/SM) and flushes the pending coefficients (<= 2 instructions per amplitude)
        const double compute = jitBaseMs30 + instrMs30 * (instr + 2.0 * rounds) + jitRoundMs30 * (rounds - 1);
        // sweep-bound launches still pay a little for their arithmetic (the overlap of HBM, FP64 and shared-memory phases is
        // not perfect): r02_s6, supremacy_30: 14 instructions per amplitude 5.75 ms, 49 -> 6.5, 76 -> 6.6, 98 -> 7.6
        return launchMs + std::ldexp(1.0, numQubits - 30) * std::max(sweepMs30() + jitUnderSweepMs30 * instr, compute);
    }
    double compute = groupBaseMs30;
    qindex targets = 0;
    for (const Gate& g : gates) {
        compute += isButterfly(g) ? gateNs[int(GateType::H)] : gateNs[int(g.type) & 31];
        if (!g.isDiagonal()) targets |= qindex(1) << g.targetQubit;
    }
    const int rounds = std::max(1, (bitCount(targets) + 3) / 4);
    compute += roundMs30 * (rounds - 1);
    // The per-gate constants are measured with every operand on a register qubit (tools/microbench.py); in groups cut from real
    // circuits controls and diagonal targets land on lane bits and rounds are uneven: the ten tile-kernel launches of
    // supremacy_30 fit  base + 1.45 * (gates + rounds)  within 12 % (profiles/r01_s16_bench_supremacy30_group.json).
    compute = groupBaseMs30 + circuitFactor * (compute - groupBaseMs30);
    return launchMs + std::ldexp(1.0, numQubits - 30) * std::max(sweepMs30(), compute);
}

unsigned long long Evaluator::signature(int numQubits) {
    loadParam(numQubits);
    unsigned long long h = 0xcbf29ce484222325ull;
    auto mix = [&](const void* p, size_t n) {
        const unsigned char* b = static_cast<const unsigned char*>(p);
        for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 0x100000001b3ull; }
    };
    const double scalars[] = {nvlinkGBs, hbmGBs, groupBaseMs30, denseBaseMs30, roundMs30, circuitFactor, launchMs, instrMs30, jitRoundMs30,
                              jitBaseMs30, jitUnderSweepMs30, specialised ? 1.0 : 0.0, fusionAware ? 1.0 : 0.0};
    mix(scalars, sizeof(scalars));
    mix(gateNs, sizeof(gateNs));
    mix(denseMs30, sizeof(denseMs30));
    return h;
}

double Evaluator::perfPerGate(int numQubits, const GateGroup* gg) { return perfPerGate(numQubits, gg->gates); }

double Evaluator::perfDense(int numQubits, const std::vector<int>& ms) {
    loadParam(numQubits);
    double compute = 0;
    compute = denseBaseMs30;
    for (int m : ms) compute += denseMs30[std::min(std::max(m, 0), 7)];
    return launchMs + std::ldexp(1.0, numQubits - 30) * std::max(sweepMs30(), compute);
}

double Evaluator::perfSwap(int numQubits, int k) {
    loadParam(numQubits);
    const double bytes = 16.0 * std::ldexp(1.0, numQubits) * (1.0 - std::ldexp(1.0, -k));
    return bytes / (nvlinkGBs * 1e9) * 1e3;
}

double Evaluator::perfBLAS(int numQubits, int blasSize) {
    loadParam(numQubits);
    return launchMs + std::ldexp(1.0, numQubits - 30) * denseMs30[std::min(std::max(blasSize, 0), 7)];
}

bool Evaluator::PerGateOrBLAS(const GateGroup* a, const GateGroup* b, int numQubits, int blasSize) {
    return perfPerGate(numQubits, a) / a->gates.size() < perfBLAS(numQubits, blasSize) / b->gates.size();
}
