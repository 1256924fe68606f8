#include "executor.h"

#include <algorithm>
#include <cassert>
#include <complex>
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
bool Executor::lowerGate(const Gate& gate, const State& state, qindex fixedMask, qindex fixedValue, hq_gate& out) {
    int localCtl[2] = {-1, -1}, nCtl = 0;
    for (int c : {gate.controlQubit, gate.controlQubit2}) {
        if (c < 0) continue;
        const int pc = state.pos[c];
        if (fixedMask >> pc & 1) {
            if (!(fixedValue >> pc & 1)) return false;
        } else {
            localCtl[nCtl++] = pc;
        }
    }
    std::memset(&out, 0, sizeof(out));
    out.control = out.control2 = -1;
    const int pt = state.pos[gate.targetQubit];
    if (fixedMask >> pt & 1) {
        if (!gate.isDiagonal()) UNREACHABLE()   // the partitioner keeps non-diagonal targets on varying bits
        const bool hi = fixedValue >> pt & 1;
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

// Fixed (non-varying) physical bits of a launch: the rank bits [L, n) always; for a per-chunk launch also the swapped local
// positions, holding the chunk number.
static void fixedBits(int numQubits, const std::vector<int>& chunkBits, int chunk, qindex& mask, qindex& value) {
    const int L = numQubits - MyGlobalVars::bit;
    mask = ((qindex(1) << numQubits) - 1) & ~((qindex(1) << L) - 1);
    value = qindex(MyMPI::rank) << L;
    for (size_t i = 0; i < chunkBits.size(); i++) {
        mask |= qindex(1) << chunkBits[i];
        if (chunk >> i & 1) value |= qindex(1) << chunkBits[i];
    }
}

static uint64_t physicalMask(const State& st, qindex logical, int numQubits) {
    uint64_t m = 0;
    for (int q = 0; q < numQubits; q++) if (logical >> q & 1) m |= 1ull << st.pos[q];
    return m;
}

static void preparePerGate(GateGroup& gg, int numQubits, const std::vector<int>& chunkBits) {
    const int L = numQubits - MyGlobalVars::bit;
    const qindex localMask = (qindex(1) << L) - 1;
    gg.tileMask = physicalMask(gg.state, gg.relatedQubits, numQubits);
    for (int chunk = 0; chunk < (1 << chunkBits.size()); chunk++) {
        qindex fmask, fvalue;
        fixedBits(numQubits, chunkBits, chunk, fmask, fvalue);
        std::vector<hq_gate> lowered;
        for (const Gate& g : gg.gates) {
            hq_gate k;
            if (Executor::lowerGate(g, gg.state, fmask, fvalue, k)) lowered.push_back(k);
        }
        hq_group_plan* plan = nullptr;
        checkHq(hq_group_plan_create_ex(L, gg.tileMask, fmask & localMask, fvalue & localMask, lowered.data(), (int)lowered.size(), &plan));
        gg.plans.push_back(plan);
    }
}

// Dense matrix of one block for the sub-state whose high index bits equal `high`: U = G_last ... G_1 acting on the
// block's qubits (matrix bit i = i-th lowest physical position).  Role of GateGroup::initCPUMatrix
// (src/schedule.cpp:578-702); gates with global controls/targets are resolved per shard by lowerGate.
static std::vector<double> buildDenseMatrix(const DenseBlock& blk, const State& state, qindex fmask, qindex fvalue,
                                            const std::vector<int>& positions) {
    const int m = (int)positions.size(), K = 1 << m;
    std::vector<std::complex<double>> U((size_t)K * K, 0.0);   // column-major: U[row + col * K]
    for (int i = 0; i < K; i++) U[(size_t)i + (size_t)i * K] = 1.0;
    int bitOf[64];
    for (int i = 0; i < 64; i++) bitOf[i] = -1;
    for (int i = 0; i < m; i++) bitOf[positions[i]] = i;
    for (const Gate& g : blk.gates) {
        hq_gate k;
        if (!Executor::lowerGate(g, state, fmask, fvalue, k)) continue;
        const std::complex<double> m00(k.mat[0], k.mat[1]), m01(k.mat[2], k.mat[3]), m10(k.mat[4], k.mat[5]), m11(k.mat[6], k.mat[7]);
        int cmask = 0;
        for (int c : {k.control, k.control2}) {
            if (c < 0) continue;
            if (bitOf[c] < 0) UNREACHABLE()
            cmask |= 1 << bitOf[c];
        }
        if (k.target < 0) {   // scalar
            for (auto& v : U) v *= m00;
            continue;
        }
        if (bitOf[k.target] < 0) UNREACHABLE()
        const int tb = 1 << bitOf[k.target];
        #pragma omp parallel for if (K >= 64 && blk.gates.size() >= 64)
        for (int col = 0; col < K; col++) {
            std::complex<double>* v = &U[(size_t)col * K];
            for (int lo = 0; lo < K; lo++) {
                if ((lo & tb) || (lo & cmask) != cmask) continue;
                const std::complex<double> a = v[lo], b = v[lo | tb];
                v[lo] = m00 * a + m01 * b;
                v[lo | tb] = m10 * a + m11 * b;
            }
        }
    }
    std::vector<double> out((size_t)K * K * 2);
    for (size_t i = 0; i < U.size(); i++) { out[2 * i] = U[i].real(); out[2 * i + 1] = U[i].imag(); }
    return out;
}

static void prepareDense(GateGroup& gg, int numQubits, const std::vector<int>& chunkBits) {
    const int L = numQubits - MyGlobalVars::bit;
    const qindex localMask = (qindex(1) << L) - 1;
    for (int chunk = 0; chunk < (1 << chunkBits.size()); chunk++) {
        qindex fmask, fvalue;
        fixedBits(numQubits, chunkBits, chunk, fmask, fvalue);
        std::vector<int> mList, qpos;
        std::vector<double> allU;
        for (const DenseBlock& blk : gg.blocks) {
            std::vector<int> positions;
            for (int q = 0; q < numQubits; q++) if (blk.qubits >> q & 1) positions.push_back(gg.state.pos[q]);
            std::sort(positions.begin(), positions.end());
            std::vector<double> U = buildDenseMatrix(blk, gg.state, fmask, fvalue, positions);
            mList.push_back((int)positions.size());
            qpos.insert(qpos.end(), positions.begin(), positions.end());
            allU.insert(allU.end(), U.begin(), U.end());
        }
        hq_dense_plan* plan = nullptr;
        checkHq(hq_dense_plan_create_ex(L, fmask & localMask, fvalue & localMask, (int)mList.size(), mList.data(), qpos.data(),
                                        allU.data(), &plan));
        gg.plans.push_back(plan);
    }
}

static void prepareGroup(GateGroup& gg, int numQubits, const std::vector<int>& chunkBits) {
    if (!gg.plans.empty()) return;
    if (gg.backend == Backend::BLAS) prepareDense(gg, numQubits, chunkBits);
    else preparePerGate(gg, numQubits, chunkBits);
}

void Executor::prepare(Schedule& schedule, int numQubits, bool hostOnly) {
    const int L = numQubits - MyGlobalVars::bit;
    for (auto& lg : schedule.localGroups) {
        const int k = (int)lg.swap.localBit.size();
        if (k > 0 && !lg.swapPlan && hostOnly == false) {
            hq_swap_plan* sp = nullptr;
            checkHq(hq_swap_plan_create(L, k, lg.swap.localBit.data(), lg.swap.globalBit.data(), &sp));
            checkHq(hq_swap_plan_set_overlap(sp, (int)lg.overlapGroups.size()));
            lg.swapPlan = sp;
        }
        for (auto& gg : lg.overlapGroups) prepareGroup(gg, numQubits, lg.swap.localBit);
        for (auto& gg : lg.fullGroups) prepareGroup(gg, numQubits, {});
    }
    if (hostOnly) return;
    // the very first launch of a run acts on |0...0>: ask for its zero-input variant (no zero fill, no read of the first sweep)
    if (zeroInputEnabled() && !schedule.localGroups.empty() && !schedule.localGroups[0].fullGroups.empty()) {
        GateGroup& first = schedule.localGroups[0].fullGroups[0];
        if (first.backend != Backend::BLAS && first.plans.size() == 1) checkHq(hq_group_plan_enable_zero_input(static_cast<hq_group_plan*>(first.plans[0])));
    }
    // tile-kernel groups run as per-group specialised kernels: fetch them from the cache, compiling the misses on all cores
    std::vector<hq_group_plan*> tilePlans;
    for (auto& lg : schedule.localGroups)
        for (auto* groups : {&lg.overlapGroups, &lg.fullGroups})
            for (auto& gg : *groups)
                if (gg.backend != Backend::BLAS)
                    for (void* p : gg.plans) tilePlans.push_back(static_cast<hq_group_plan*>(p));
    checkHq(hq_group_plans_warm(tilePlans.data(), (int)tilePlans.size()));
}

void Executor::release(Schedule& schedule) {
    for (auto& lg : schedule.localGroups) {
        if (lg.swapPlan) hq_swap_plan_destroy(static_cast<hq_swap_plan*>(lg.swapPlan));
        lg.swapPlan = nullptr;
        for (auto* groups : {&lg.overlapGroups, &lg.fullGroups})
            for (auto& gg : *groups) {
                for (void* p : gg.plans) {
                    if (gg.backend == Backend::BLAS) hq_dense_plan_destroy(static_cast<hq_dense_plan*>(p));
                    else hq_group_plan_destroy(static_cast<hq_group_plan*>(p));
                }
                gg.plans.clear();
            }
    }
}

bool Executor::zeroInputEnabled() {
    static const bool on = !(getenv("HQ_ZERO_INPUT") && atoi(getenv("HQ_ZERO_INPUT")) == 0);
    return on;
}

bool Executor::runFromZero() {
    if (!zeroInputEnabled() || schedule.localGroups.empty() || schedule.localGroups[0].fullGroups.empty()) return false;
    GateGroup& first = schedule.localGroups[0].fullGroups[0];
    if (first.backend == Backend::BLAS || first.plans.size() != 1) return false;
    if (perGroupMs) checkHq(hq_timer_start());
    const int rc = hq_group_plan_launch_from_zero(static_cast<hq_group_plan*>(first.plans[0]), deviceStateVec[0], MyMPI::rank == 0);
    if (rc == HQ_ERR_UNSUPPORTED) return false;   // nothing was launched
    checkHq(rc);
    if (perGroupMs) {
        float ms = 0;
        checkHq(hq_timer_stop_ms(&ms));
        perGroupMs->push_back(ms);
    }
    firstFromZero = true;
    run();
    return true;
}

void Executor::applyGateGroup(GateGroup& gg, int chunk) {
    if (firstFromZero && &gg == &schedule.localGroups[0].fullGroups[0]) { firstFromZero = false; return; }   // already done by runFromZero
    if (perGroupMs) checkHq(hq_timer_start());
    auto launch = [&](void* plan, qComplex* base) {
        if (gg.backend == Backend::BLAS) checkHq(hq_dense_plan_launch(static_cast<hq_dense_plan*>(plan), base, 0))
        else checkHq(hq_group_plan_launch(static_cast<hq_group_plan*>(plan), base, 0))
    };
    launch(gg.plans[chunk < 0 ? 0 : chunk], deviceStateVec[0]);   // a per-chunk plan carries its chunk's fixed bits
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
