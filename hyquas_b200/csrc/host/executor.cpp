#include "executor.h"

#include <cassert>
#include <cstring>

#include "logger.h"
#include "swap.h"

Executor::Executor(std::vector<qComplex*> deviceStateVec_, int numQubits_, Schedule& schedule_)
    : deviceStateVec(std::move(deviceStateVec_)), numQubits(numQubits_), schedule(schedule_) {}

// Semantics follow the reference's Executor::getGate (src/executor.cpp:190-403) but as one rule instead of a
// case table:  a control on a non-local bit either removes the gate (bit = 0) or disappears (bit = 1);
// a (necessarily diagonal) target on a non-local bit selects d = m00 or m11 and the gate degenerates to
// "multiply by d where the remaining local controls are 1": a scalar (GCC/GZZ/GII there), a one-qubit
// diag(1,d) on the control (Z/U1/GOC there), or a controlled diag(1,d).
bool Executor::lowerGate(const Gate& gate, const State& state, int numLocal, qindex highIndex, hq_gate& out) {
    int localCtl[2] = {-1, -1}, nCtl = 0;
    for (int c : {gate.controlQubit, gate.controlQubit2}) {
        if (c < 0) continue;
        const int pc = state.pos[c];
        if (pc >= numLocal) {
            if (!(highIndex >> (pc - numLocal) & 1)) return false;
        } else {
            localCtl[nCtl++] = pc;
        }
    }
    std::memset(&out, 0, sizeof(out));
    out.control = out.control2 = -1;
    const int pt = state.pos[gate.targetQubit];
    if (pt >= numLocal) {
        if (!gate.isDiagonal()) UNREACHABLE()   // the partitioner keeps non-diagonal targets local
        const bool hi = highIndex >> (pt - numLocal) & 1;
        const qComplex d = gate.mat[hi][hi];
        if (d.x == 1.0 && d.y == 0.0) return false;
        if (nCtl == 0) {
            out.type = HQ_GCC; out.target = -1;
            out.mat[0] = d.x; out.mat[1] = d.y; out.mat[6] = d.x; out.mat[7] = d.y;
        } else {
            out.type = HQ_GOC; out.target = localCtl[0]; out.control = localCtl[1];
            out.mat[0] = 1.0; out.mat[6] = d.x; out.mat[7] = d.y;
        }
        return true;
    }
    out.type = (int)gate.type;
    out.target = pt;
    out.control = localCtl[0];
    out.control2 = localCtl[1];
    for (int i = 0; i < 4; i++) {
        out.mat[2 * i] = gate.mat[i >> 1][i & 1].x;
        out.mat[2 * i + 1] = gate.mat[i >> 1][i & 1].y;
    }
    return true;
}

static uint64_t physicalMask(const State& st, qindex logical, int numQubits) {
    uint64_t m = 0;
    for (int q = 0; q < numQubits; q++) if (logical >> q & 1) m |= 1ull << st.pos[q];
    return m;
}

static void preparePerGate(GateGroup& gg, int numQubits, int numLocal, int numChunks) {
    const qindex rank = MyMPI::rank;
    gg.tileMask = physicalMask(gg.state, gg.relatedQubits, numQubits);
    for (int chunk = 0; chunk < numChunks; chunk++) {
        const qindex high = numChunks > 1 ? ((rank * numChunks) | chunk) : rank;
        std::vector<hq_gate> lowered;
        for (const Gate& g : gg.gates) {
            hq_gate k;
            if (Executor::lowerGate(g, gg.state, numLocal, high, k)) lowered.push_back(k);
        }
        hq_group_plan* plan = nullptr;
        checkHq(hq_group_plan_create(numLocal, gg.tileMask, lowered.data(), (int)lowered.size(), &plan));
        gg.plans.push_back(plan);
    }
}

void Executor::prepare(Schedule& schedule, int numQubits, bool hostOnly) {
    const int L = numQubits - MyGlobalVars::bit;
    for (auto& lg : schedule.localGroups) {
        const int k = (int)lg.swap.localBit.size();
        if (k > 0 && !lg.swapPlan && hostOnly == false) {
            hq_swap_plan* sp = nullptr;
            checkHq(hq_swap_plan_create(L, k, lg.swap.localBit.data(), lg.swap.globalBit.data(), &sp));
            lg.swapPlan = sp;
        }
        for (auto& gg : lg.overlapGroups) if (gg.plans.empty()) preparePerGate(gg, numQubits, L - k, 1 << k);
        for (auto& gg : lg.fullGroups) if (gg.plans.empty()) preparePerGate(gg, numQubits, L, 1);
    }
}

void Executor::release(Schedule& schedule) {
    for (auto& lg : schedule.localGroups) {
        if (lg.swapPlan) hq_swap_plan_destroy(static_cast<hq_swap_plan*>(lg.swapPlan));
        lg.swapPlan = nullptr;
        for (auto* groups : {&lg.overlapGroups, &lg.fullGroups})
            for (auto& gg : *groups) {
                for (void* p : gg.plans) hq_group_plan_destroy(static_cast<hq_group_plan*>(p));
                gg.plans.clear();
            }
    }
}

void Executor::applyGateGroup(GateGroup& gg, int chunk) {
    const int L = numQubits - MyGlobalVars::bit;
    if (perGroupMs) checkHq(hq_timer_start());
    if (chunk < 0) {
        checkHq(hq_group_plan_launch(static_cast<hq_group_plan*>(gg.plans[0]), deviceStateVec[0], 0));
    } else {
        const int nChunks = (int)gg.plans.size();
        int k = 0;
        while ((1 << k) < nChunks) k++;
        qComplex* base = deviceStateVec[0] + ((qindex)chunk << (L - k));
        checkHq(hq_group_plan_launch(static_cast<hq_group_plan*>(gg.plans[chunk]), base, 0));
    }
    if (perGroupMs) {
        float ms = 0;
        checkHq(hq_timer_stop_ms(&ms));
        perGroupMs->push_back(ms);
    }
}

void Executor::run() {
    for (size_t s = 0; s < schedule.localGroups.size(); s++) {
        LocalGroup& lg = schedule.localGroups[s];
        if (s > 0 && !lg.swap.empty()) {
            hyquas::SwapExec ex(deviceStateVec[0], numQubits - MyGlobalVars::bit, lg.swap, lg.swapPlan);
            const int nChunks = 1 << lg.swap.localBit.size();
            ex.begin();
            for (int i = 0; i < nChunks; i++) {
                const int chunk = ex.waitNextChunk();   // compute stream now waits for this chunk's arrival
                for (auto& gg : lg.overlapGroups) applyGateGroup(gg, chunk);
            }
            ex.end();
        }
        for (auto& gg : lg.fullGroups) applyGateGroup(gg, -1);
    }
    finalize();
}

void Executor::finalize() {
    checkHq(hq_sync());
    schedule.finalState = schedule.localGroups.empty() ? State(numQubits) : schedule.localGroups.back().state;
}
