// Circuit-level C-ABI (declared in include/hyquas_b200_circuit.h): lets non-C++ hosts (the ctypes tests,
// bench.py) drive exactly the code path `hyquas_main` runs: parse / addGate -> compile -> run -> dump.
#include <algorithm>
#include <cstring>
#include <memory>

#include "circuit.h"
#include "hyquas_b200_circuit.h"
#include "logger.h"
#include "qasm.h"
#include "compiler.h"
#include "peephole.h"

struct hq_circuit {
    std::unique_ptr<Circuit> c;
    std::string dump;
};

static thread_local std::string g_cerr;
extern "C" const char* hq_circuit_last_error(void) { return g_cerr.c_str(); }

extern "C" int hq_runtime_init(void) {
    static bool done = false;
    if (!done) { MyGlobalVars::init(); done = true; }
    return HQ_OK;
}

extern "C" int hq_runtime_init_host_only(int world_size, int rank) {
    MyGlobalVars::initForTest(world_size, rank);
    return HQ_OK;
}

extern "C" int hq_circuit_create(int num_qubits, hq_circuit** out) {
    if (!out || num_qubits < 1 || num_qubits > 40) { g_cerr = "bad qubit count"; return HQ_ERR_ARG; }
    *out = new hq_circuit{std::unique_ptr<Circuit>(new Circuit(num_qubits)), ""};
    return HQ_OK;
}

extern "C" int hq_circuit_from_qasm(const char* text, hq_circuit** out) {
    if (!text || !out) { g_cerr = "null argument"; return HQ_ERR_ARG; }
    std::string err;
    auto c = hyquas::parseQasmText(text, err);
    if (!c) { g_cerr = err; return HQ_ERR_ARG; }
    *out = new hq_circuit{std::move(c), ""};
    return HQ_OK;
}

extern "C" int hq_circuit_add_gate(hq_circuit* h, int type, int control2, int control, int target, const double* params, int nparams) {
    if (!h) { g_cerr = "null circuit"; return HQ_ERR_ARG; }
    auto P = [&](int i) { return i < nparams && params ? params[i] : 0.0; };
    const int n = h->c->numQubits;
    for (int q : {control2, control}) if (q < -1 || q >= n) { g_cerr = "control out of range"; return HQ_ERR_ARG; }
    if (target < 0 || target >= n) { g_cerr = "target out of range"; return HQ_ERR_ARG; }
    {   // operands must be distinct, and a controlled type needs its control(s)
        const GateType t = (GateType)type;
        const bool two = t == GateType::CCX;
        const bool one = t == GateType::CNOT || t == GateType::CY || t == GateType::CZ || t == GateType::CRX || t == GateType::CRY ||
                         t == GateType::CU1 || t == GateType::CRZ;
        if ((one || two) && control < 0) { g_cerr = "controlled gate without a control qubit"; return HQ_ERR_ARG; }
        if (two && control2 < 0) { g_cerr = "ccx needs two control qubits"; return HQ_ERR_ARG; }
        if (!one && !two && (control >= 0 || control2 >= 0)) { g_cerr = "control qubit given to an uncontrolled gate type"; return HQ_ERR_ARG; }
        if (!two && control2 >= 0) { g_cerr = "second control given to a singly controlled gate type"; return HQ_ERR_ARG; }
        if ((control >= 0 && control == target) || (control2 >= 0 && (control2 == target || control2 == control))) {
            g_cerr = "gate operands must be distinct qubits";
            return HQ_ERR_ARG;
        }
    }
    Gate g;
    switch ((GateType)type) {
        case GateType::CCX: g = Gate::CCX(control, control2, target); break;
        case GateType::CNOT: g = Gate::CNOT(control, target); break;
        case GateType::CY: g = Gate::CY(control, target); break;
        case GateType::CZ: g = Gate::CZ(control, target); break;
        case GateType::CRX: g = Gate::CRX(control, target, P(0)); break;
        case GateType::CRY: g = Gate::CRY(control, target, P(0)); break;
        case GateType::CU1: g = Gate::CU1(control, target, P(0)); break;
        case GateType::CRZ: g = Gate::CRZ(control, target, P(0)); break;
        case GateType::U1: g = Gate::U1(target, P(0)); break;
        case GateType::U2: g = Gate::U2(target, P(0), P(1)); break;
        case GateType::U3: g = Gate::U3(target, P(0), P(1), P(2)); break;
        case GateType::H: g = Gate::H(target); break;
        case GateType::X: g = Gate::X(target); break;
        case GateType::Y: g = Gate::Y(target); break;
        case GateType::Z: g = Gate::Z(target); break;
        case GateType::S: g = Gate::S(target); break;
        case GateType::SDG: g = Gate::SDG(target); break;
        case GateType::T: g = Gate::T(target); break;
        case GateType::TDG: g = Gate::TDG(target); break;
        case GateType::RX: g = Gate::RX(target, P(0)); break;
        case GateType::RY: g = Gate::RY(target, P(0)); break;
        case GateType::RZ: g = Gate::RZ(target, P(0)); break;
        default: g_cerr = "unsupported gate type"; return HQ_ERR_ARG;
    }
    h->c->addGate(g);
    return HQ_OK;
}

extern "C" int hq_circuit_num_qubits(const hq_circuit* h) { return h ? h->c->numQubits : -1; }
extern "C" int hq_circuit_num_gates(const hq_circuit* h) { return h ? (int)h->c->getGates().size() : -1; }

extern "C" int hq_circuit_compile(hq_circuit* h) {
    if (!h) { g_cerr = "null circuit"; return HQ_ERR_ARG; }
    g_cerr = h->c->compileError();   // what the CLI reports with exit(1) is an error code for an embedding host
    if (!g_cerr.empty()) return HQ_ERR_ARG;
    h->c->compile();
    return HQ_OK;
}

// host-only: run the partitioner without touching a GPU and report the shape of the schedule
extern "C" int hq_circuit_plan_only(hq_circuit* h, int* stages, int* groups, int* swapped_bits) {
    if (!h) { g_cerr = "null circuit"; return HQ_ERR_ARG; }
    g_cerr = h->c->compileError();
    if (!g_cerr.empty()) return HQ_ERR_ARG;
    Compiler compiler(h->c->numQubits, hyquas::peephole(h->c->getGates(), nullptr));   // what compile() partitions
    Schedule s = compiler.run();
    if (stages) *stages = (int)s.localGroups.size();
    if (groups) *groups = s.numGroups();
    if (swapped_bits) { *swapped_bits = 0; for (auto& lg : s.localGroups) *swapped_bits += (int)lg.swap.localBit.size(); }
    return HQ_OK;
}

// Test hook: partition the circuit as compile() would and check the schedule against the gate list (hyquas::checkSchedule);
// HQ_OK and an empty message when it is a valid reordering.  Works without a GPU (plan only, nothing is lowered or uploaded).
extern "C" int hq_debug_schedule_check(hq_circuit* h, char* why, size_t cap) {
    if (!h) { g_cerr = "null circuit"; return HQ_ERR_ARG; }
    g_cerr = h->c->compileError();
    if (!g_cerr.empty()) return HQ_ERR_ARG;
    const std::vector<Gate> gates = hyquas::peephole(h->c->getGates(), nullptr);
    Compiler compiler(h->c->numQubits, gates);
    Schedule s = compiler.run();
    if (const char* e = getenv("HQ_TEST_BREAK_SCHEDULE")) {   // (the tests of the checker itself: damage the schedule first)
        std::vector<GateGroup>& first = s.localGroups.front().fullGroups;
        std::vector<GateGroup>& last = s.localGroups.back().fullGroups;
        if (!first.empty() && !last.empty() && !last.back().gates.empty()) {
            if (atoi(e) == 1) first.front().gates.insert(first.front().gates.begin(), last.back().gates.back());   // a gate twice
            else if (atoi(e) == 2) last.back().gates.pop_back();                                                   // a gate missing
            else if (atoi(e) == 3) std::swap(first.front().gates, last.back().gates);                              // tiles
            else std::reverse(first.front().gates.begin(), first.front().gates.end());                             // order
        }
    }
    std::vector<Gate> numbered = gates;
    for (size_t i = 0; i < numbered.size(); i++) numbered[i].gateID = (int)i;   // the ids the Compiler gave its copy
    const std::string bad = hyquas::checkSchedule(s, numbered, h->c->numQubits, h->c->numQubits - MyGlobalVars::bit, compiler.tileBits);
    if (why && cap) snprintf(why, cap, "%s", bad.c_str());
    if (!bad.empty()) { g_cerr = bad; return HQ_ERR_UNSUPPORTED; }
    return HQ_OK;
}

extern "C" int hq_circuit_run(hq_circuit* h, int copy_back, int destroy, int* time_us, double* device_ms) {
    if (!h) { g_cerr = "null circuit"; return HQ_ERR_ARG; }
    const int us = h->c->run(copy_back != 0, destroy != 0);
    if (time_us) *time_us = us;
    if (device_ms) *device_ms = h->c->lastDeviceMs;
    return HQ_OK;
}

extern "C" int hq_circuit_prepare_state(hq_circuit* h) {
    if (!h) { g_cerr = "null circuit"; return HQ_ERR_ARG; }
    h->c->prepareState();
    return HQ_OK;
}

extern "C" int hq_circuit_execute(hq_circuit* h, int* time_us, double* device_ms, float* per_group_ms, int cap, int* ngroups) {
    if (!h) { g_cerr = "null circuit"; return HQ_ERR_ARG; }
    std::vector<float> per;
    const int us = h->c->execute(per_group_ms ? &per : nullptr);
    if (time_us) *time_us = us;
    if (device_ms) *device_ms = h->c->lastDeviceMs;
    if (per_group_ms) for (int i = 0; i < cap && i < (int)per.size(); i++) per_group_ms[i] = per[i];
    if (ngroups) *ngroups = (int)per.size();
    return HQ_OK;
}

// Frees the resident state vector (run(destroy = false) / prepare_state keep it); the compiled schedule and its plans stay.
extern "C" int hq_circuit_release_state(hq_circuit* h) {
    if (!h) { g_cerr = "null circuit"; return HQ_ERR_ARG; }
    h->c->destroyState();
    return HQ_OK;
}

extern "C" int hq_circuit_measure(hq_circuit* h, int qubit, double* p0) {
    if (!h || !p0) { g_cerr = "null argument"; return HQ_ERR_ARG; }
    if (qubit < 0 || qubit >= h->c->numQubits) { g_cerr = "qubit out of range"; return HQ_ERR_ARG; }
    *p0 = h->c->measure(qubit);
    return HQ_OK;
}

extern "C" int hq_circuit_norm2(hq_circuit* h, double* out) {
    if (!h || !out) { g_cerr = "null argument"; return HQ_ERR_ARG; }
    *out = h->c->norm2();
    return HQ_OK;
}

// Circuit::ampAt (src/circuit.cpp:95-125): collective when there is more than one process -- every rank must call it
extern "C" int hq_circuit_amp_at(hq_circuit* h, long long idx, double out_re_im[2]) {
    if (!h || !out_re_im) { g_cerr = "null argument"; return HQ_ERR_ARG; }
    if (idx < 0 || idx >= (1ll << h->c->numQubits)) { g_cerr = "amplitude index out of range"; return HQ_ERR_ARG; }
    const ResultItem it = h->c->ampAt(idx);
    out_re_im[0] = it.amp.x;
    out_re_im[1] = it.amp.y;
    return HQ_OK;
}

extern "C" int hq_circuit_swap_alone_ms(hq_circuit* h, double* ms) {
    if (!h || !ms) { g_cerr = "null argument"; return HQ_ERR_ARG; }
    *ms = h->c->swapAloneMs();
    return HQ_OK;
}

extern "C" int hq_circuit_io_bytes(const hq_circuit* h, size_t* h2d_plan_bytes, size_t* d2h_dump_bytes) {
    if (!h) { g_cerr = "null circuit"; return HQ_ERR_ARG; }
    if (h2d_plan_bytes) *h2d_plan_bytes = h->c->planBytes();
    if (d2h_dump_bytes) *d2h_dump_bytes = h->c->dumpBytes();
    return HQ_OK;
}

extern "C" int hq_circuit_schedule_info(const hq_circuit* h, int* stages, int* groups, int* gates_in_groups) {
    if (!h) { g_cerr = "null circuit"; return HQ_ERR_ARG; }
    const Schedule& s = h->c->getSchedule();
    if (stages) *stages = (int)s.localGroups.size();
    if (groups) *groups = s.numGroups();
    if (gates_in_groups) {
        int g = 0;
        for (auto& lg : s.localGroups) {
            for (auto& gg : lg.fullGroups) g += (int)gg.gates.size();
            for (auto& gg : lg.overlapGroups) g += (int)gg.gates.size();
        }
        *gates_in_groups = g;
    }
    return HQ_OK;
}

// Group i in execution order (overlap groups of a stage first, then its full groups): backend 1 = tile kernel,
// 2 = fused dense kernel; launches = kernel launches it costs (2^k for a per-chunk group); predicted_ms = evaluator.
extern "C" int hq_circuit_group_info(const hq_circuit* h, int index, int* backend, int* ngates, double* predicted_ms, int* launches,
                                     int* nblocks) {
    if (!h) { g_cerr = "null circuit"; return HQ_ERR_ARG; }
    int i = 0;
    for (const auto& lg : h->c->getSchedule().localGroups) {
        const int k = (int)lg.swap.localBit.size();
        for (const auto* groups : {&lg.overlapGroups, &lg.fullGroups})
            for (const auto& gg : *groups) {
                if (i++ != index) continue;
                const bool chunked = groups == &lg.overlapGroups;
                if (backend) *backend = gg.backend == Backend::BLAS ? 2 : 1;
                if (ngates) *ngates = (int)gg.gates.size();
                if (predicted_ms) *predicted_ms = gg.predictedMs * (chunked ? (1 << k) : 1);
                if (launches) *launches = chunked ? (1 << k) : 1;
                if (nblocks) *nblocks = (int)gg.blocks.size();
                return HQ_OK;
            }
    }
    g_cerr = "group index out of range";
    return HQ_ERR_ARG;
}

// Tile-kernel group `index` (execution order, as hq_circuit_group_info): register rounds and FP64 instructions per amplitude of its
// (first) plan; -1 / -1 for dense groups.
extern "C" int hq_circuit_group_cost(const hq_circuit* h, int index, int* rounds, double* fp64_per_amp) {
    if (!h) { g_cerr = "null circuit"; return HQ_ERR_ARG; }
    int i = 0;
    for (const auto& lg : h->c->getSchedule().localGroups)
        for (const auto* groups : {&lg.overlapGroups, &lg.fullGroups})
            for (const auto& gg : *groups) {
                if (i++ != index) continue;
                if (rounds) *rounds = -1;
                if (fp64_per_amp) *fp64_per_amp = -1;
                if (gg.backend != Backend::BLAS && !gg.plans.empty()) return hq_group_plan_cost(static_cast<const hq_group_plan*>(gg.plans[0]), rounds, fp64_per_amp);
                return HQ_OK;
            }
    g_cerr = "group index out of range";
    return HQ_ERR_ARG;
}

extern "C" int hq_circuit_dump(hq_circuit* h, char* buf, size_t cap, size_t* needed) {
    if (!h) { g_cerr = "null circuit"; return HQ_ERR_ARG; }
    h->dump = h->c->stateDump();
    if (needed) *needed = h->dump.size() + 1;
    if (buf && cap) {
        const size_t n = std::min(cap - 1, h->dump.size());
        memcpy(buf, h->dump.data(), n);
        buf[n] = 0;
    }
    return HQ_OK;
}

extern "C" int hq_circuit_amplitudes(hq_circuit* h, double* out_re_im) {
    if (!h || !out_re_im) { g_cerr = "null argument"; return HQ_ERR_ARG; }
    std::vector<qComplex> st;
    if (!h->c->fullState(st)) { g_cerr = "full state not available (run with destroy=0 or copy_back=1, single GPU, n<=30)"; return HQ_ERR_UNSUPPORTED; }
    memcpy(out_re_im, st.data(), st.size() * sizeof(qComplex));
    return HQ_OK;
}

// This process' shard in PHYSICAL order (2^(n - log2 world) amplitudes) and the final layout pos[logical] = physical:
// what a multi-process host needs to assemble or spot-check the distributed state (Circuit::run's copy_back +
// toPhysicalID, src/circuit.cpp:54-63,95-104).
extern "C" int hq_circuit_local_shard(hq_circuit* h, double* out_re_im) {
    if (!h || !out_re_im) { g_cerr = "null argument"; return HQ_ERR_ARG; }
    if (!h->c->localShard(out_re_im)) { g_cerr = "state not resident (run with destroy=0)"; return HQ_ERR_UNSUPPORTED; }
    return HQ_OK;
}

extern "C" int hq_circuit_final_layout(hq_circuit* h, int* pos) {
    if (!h || !pos) { g_cerr = "null argument"; return HQ_ERR_ARG; }
    const State& st = h->c->getSchedule().finalState;
    for (int i = 0; i < h->c->numQubits; i++) pos[i] = st.pos[i];
    return HQ_OK;
}

extern "C" int hq_circuit_logger_flush(char* buf, size_t cap) {
    std::string all;
    for (const auto& s : Logger::pending()) all += "Logger: " + s + "\n";
    if (buf && cap) {
        const size_t n = std::min(cap - 1, all.size());
        memcpy(buf, all.data(), n);
        buf[n] = 0;
    }
    Logger::print();   // clears; also echoes to stdout like the CLI
    return HQ_OK;
}

extern "C" int hq_circuit_destroy(hq_circuit* h) {
    delete h;
    return HQ_OK;
}

// ---- TEST HOOK (CPU test-suite only; never used by run()) ---------------------------------------------------
// Replays the compiled schedule of a single-process circuit on a host array by interpreting each group's device
// plan with the plan emulator (device/plan_emulator.cpp).  Validates partitioner + lowering + round planner
// without a GPU; the product path launches the CUDA kernels instead.
extern "C" int hq_debug_group_plan_emulate(const hq_group_plan* plan, double* state_re_im);
extern "C" int hq_debug_dense_plan_emulate(const hq_dense_plan* plan, double* state_re_im);
static int emulateGroup(const GateGroup& gg, int idx, double* state) {
    if (gg.backend == Backend::BLAS) return hq_debug_dense_plan_emulate(static_cast<const hq_dense_plan*>(gg.plans.at(idx)), state);
    return hq_debug_group_plan_emulate(static_cast<const hq_group_plan*>(gg.plans.at(idx)), state);
}
// Stepwise variant for the multi-process CPU tests (world_size-2 gloo): the test moves the data between ranks itself,
// following the stage's SwapPlan, and asks this hook to replay the stage's gate groups on its shard.
//   hq_debug_stage_swap:    the swap that establishes `stage` (npairs local bit transpositions, then k (local, global) trades)
//   hq_debug_stage_emulate: phase 0 = overlap groups on chunk `chunk` (pointer = start of the shard), phase 1 = full groups
//   hq_debug_final_pos:     pos[logical qubit] = physical bit after the last stage
extern "C" int hq_debug_num_stages(hq_circuit* h) { return h ? (int)h->c->getSchedule().localGroups.size() : -1; }
extern "C" int hq_debug_stage_swap(hq_circuit* h, int stage, int* npairs, int* pa, int* pb, int* k, int* lbits, int* gbits,
                                   int* noverlap) {
    if (!h || stage < 0 || stage >= (int)h->c->getSchedule().localGroups.size()) { g_cerr = "bad stage"; return HQ_ERR_ARG; }
    const LocalGroup& lg = h->c->getSchedule().localGroups[stage];
    *npairs = (int)lg.swap.localPerm.size();
    for (int i = 0; i < *npairs; i++) { pa[i] = lg.swap.localPerm[i].first; pb[i] = lg.swap.localPerm[i].second; }
    *k = (int)lg.swap.localBit.size();
    for (int i = 0; i < *k; i++) { lbits[i] = lg.swap.localBit[i]; gbits[i] = lg.swap.globalBit[i]; }
    if (noverlap) *noverlap = (int)lg.overlapGroups.size();
    return HQ_OK;
}
extern "C" int hq_debug_stage_emulate(hq_circuit* h, int stage, int phase, int chunk, double* state_re_im) {
    if (!h || !state_re_im || stage < 0 || stage >= (int)h->c->getSchedule().localGroups.size()) { g_cerr = "bad stage"; return HQ_ERR_ARG; }
    const LocalGroup& lg = h->c->getSchedule().localGroups[stage];
    if (phase == 0) {
        for (const auto& gg : lg.overlapGroups) {   // a per-chunk plan carries its chunk's fixed bits: whole shard pointer
            int rc = emulateGroup(gg, chunk, state_re_im);
            if (rc != HQ_OK) return rc;
        }
    } else {
        for (const auto& gg : lg.fullGroups) {
            int rc = emulateGroup(gg, 0, state_re_im);
            if (rc != HQ_OK) return rc;
        }
    }
    return HQ_OK;
}
extern "C" int hq_debug_final_pos(hq_circuit* h, int* pos) {
    if (!h || !pos) { g_cerr = "null argument"; return HQ_ERR_ARG; }
    const State& st = h->c->getSchedule().finalState;
    for (int i = 0; i < h->c->numQubits; i++) pos[i] = st.pos[i];
    return HQ_OK;
}

extern "C" int hq_debug_circuit_emulate(hq_circuit* h, double* state_re_im) {
    if (!h || !state_re_im) { g_cerr = "null argument"; return HQ_ERR_ARG; }
    if (MyGlobalVars::numGPUs != 1) { g_cerr = "emulation hook is single-process"; return HQ_ERR_UNSUPPORTED; }
    for (const auto& lg : h->c->getSchedule().localGroups)
        for (const auto& gg : lg.fullGroups) {
            int rc = emulateGroup(gg, 0, state_re_im);
            if (rc != HQ_OK) return rc;
        }
    return HQ_OK;
}
