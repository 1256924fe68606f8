#include "evaluator.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>

Evaluator* Evaluator::getInstance() {
    static Evaluator* inst = new Evaluator();
    return inst;
}

Evaluator::Evaluator() {
    // Defaults: B200, FP64.  In-register cost per gate class = FP64 issue slots per amplitude / (148 SMs x 64
    // FP64 lanes x ~1.7 GHz sustained), expressed in ms per 2^30 amplitudes; refined by tools/calibrate.py.
    hbmGBs = 5800.0;
    launchMs = 0.01;
    nvlinkGBs = 700.0;
    roundMs30 = 0.55;
    const double slot = 1073741824.0 / (148.0 * 64.0 * 1.7e9) * 1e3;   // ms per FP64 slot per amplitude at 2^30
    for (auto& g : gateNs) g = 4 * slot;
    auto set = [&](GateType t, double slots) { gateNs[int(t)] = slots * slot; };
    set(GateType::CCX, 0.3); set(GateType::CNOT, 0.5); set(GateType::X, 1.0);
    set(GateType::CY, 1.0); set(GateType::Y, 2.0);
    set(GateType::CZ, 0.6); set(GateType::Z, 1.2);
    set(GateType::CRX, 2.2); set(GateType::CRY, 2.2); set(GateType::RX, 4.2); set(GateType::RY, 4.2);
    set(GateType::CU1, 1.2); set(GateType::CRZ, 2.4); set(GateType::U1, 2.2); set(GateType::RZ, 4.2);
    set(GateType::U2, 8.5); set(GateType::U3, 8.5); set(GateType::H, 4.2);
    set(GateType::S, 2.2); set(GateType::SDG, 2.2); set(GateType::T, 2.2); set(GateType::TDG, 2.2);
    for (int m = 0; m < 8; m++) denseMs30[m] = std::max(32.0 * 1073741824.0 / (hbmGBs * 1e9) * 1e3, 4.0 * (1 << m) * slot);
}

void Evaluator::loadParam(int) {
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
        else if (key == "gate") { int i; double v; in >> i >> v; if (i >= 0 && i < 32) gateNs[i] = v; }
        else if (key == "dense") { int i; double v; in >> i >> v; if (i >= 0 && i < 8) denseMs30[i] = v; }
    }
}

double Evaluator::perfPerGate(int numQubits, const std::vector<GateType>& types) {
    loadParam(numQubits);
    const double scale = std::ldexp(1.0, numQubits - 30);
    double compute = 0;
    for (GateType t : types) compute += gateNs[int(t) & 31];
    compute += roundMs30 * (1 + types.size() / 24.0);
    const double sweep = 32.0 * 1073741824.0 / (hbmGBs * 1e9) * 1e3;
    return launchMs + scale * std::max(sweep, compute);
}

double Evaluator::perfPerGate(int numQubits, const GateGroup* gg) {
    std::vector<GateType> tys;
    for (const Gate& g : gg->gates) tys.push_back(g.type);
    return perfPerGate(numQubits, tys);
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
