#include "circuit.h"
#include "peephole.h"

#include <algorithm>
#include <cassert>
#include <chrono>

#include "compiler.h"
#include "executor.h"
#include "logger.h"
#include "swap.h"
using namespace std;

std::string ResultItem::str() const {   // "%lld %.12f: %.12f %.12f\n" = idx |amp|^2 : re im (reference circuit.h:14-16)
    char buf[160];
    snprintf(buf, sizeof(buf), "%lld %.12f: %.12f %.12f\n", idx, amp.x * amp.x + amp.y * amp.y, zero_wrapper(amp.x), zero_wrapper(amp.y));
    return buf;
}

Circuit::~Circuit() {
    Executor::release(schedule);
    destroyState();
}

void Circuit::destroyState() {
    if (!deviceStateVec.empty() && MyGlobalVars::numGPUs > 1 && !MyGlobalVars::hostOnly) checkHq(hq_swap_detach());
    for (auto* p : deviceStateVec) hq_state_free(p);
    deviceStateVec.clear();
}

std::string Circuit::compileError() const {
    // The tile kernel works on tiles of >= 2^10 amplitudes (the reference's LOCAL_QUBIT_SIZE = 10, utils.h:45).  One process
    // can fall back to the dense kernel for 9 local qubits; with global qubits a group may hold gates whose controls sit on
    // global or fixed bits, which only the tile kernel resolves, so 10 local qubits are required there.
    const int L = numQubits - MyGlobalVars::bit;
    const int need = MyGlobalVars::bit > 0 ? 10 : 9;
    if (L >= need) return "";
    char buf[200];
    snprintf(buf, sizeof(buf), "hyquas_b200: %d qubits on %d GPU(s) leaves %d local qubits; at least %d are needed", numQubits,
             MyGlobalVars::numGPUs, L, need);
    return buf;
}

void Circuit::compile() {
    const std::string bad = compileError();
    if (!bad.empty()) {
        printf("%s\n", bad.c_str());
        exit(1);
    }
    auto t0 = chrono::system_clock::now();
    Logger::add("Total Gates %d", int(gates.size()));
    Executor::release(schedule);
    hyquas::PeepholeStats ph;
    const std::vector<Gate> optimised = hyquas::peephole(gates, &ph);
    if (ph.zzPatterns + ph.hcxhPatterns + ph.xdxPatterns + ph.mergedPairs > 0)
        Logger::add("Peephole: %d -> %d gates (%d cx-diag-cx, %d h-cx-h, %d x-diag-x patterns made diagonal; %d single-qubit pairs merged)",
                    ph.gatesIn, ph.gatesOut, ph.zzPatterns, ph.hcxhPatterns, ph.xdxPatterns, ph.mergedPairs);
    Compiler compiler(numQubits, optimised);
    schedule = compiler.run();
    int fullGroups = 0, fullGates = 0, overlapGates = 0;
    for (auto& lg : schedule.localGroups) {
        fullGroups += (int)lg.fullGroups.size();
        for (auto& gg : lg.fullGroups) fullGates += (int)gg.gates.size();
        for (auto& gg : lg.overlapGroups) overlapGates += (int)gg.gates.size();
    }
    Logger::add("Total Groups: %d %d %d %d", int(schedule.localGroups.size()), fullGroups, fullGates, overlapGates);
    if (getenv("HQ_SHOW_SCHEDULE")) schedule.dump(numQubits);
    auto t1 = chrono::system_clock::now();
    Executor::prepare(schedule, numQubits, MyGlobalVars::hostOnly);
    auto t2 = chrono::system_clock::now();
    const int d1 = (int)chrono::duration_cast<chrono::microseconds>(t1 - t0).count();
    const int d2 = (int)chrono::duration_cast<chrono::microseconds>(t2 - t1).count();
    Logger::add("Compile Time: %d us + %d us = %d us", d1, d2, d1 + d2);
    compiled = true;
}

void Circuit::allocState() {
    if (!compiled) compile();
    const int L = numQubits - MyGlobalVars::bit;
    if (deviceStateVec.empty()) {
        void* st = nullptr;
        checkHq(hq_state_alloc(L, &st));
        deviceStateVec.assign(1, static_cast<qComplex*>(st));
        if (MyGlobalVars::numGPUs > 1) checkHq(hq_swap_attach(st));   // p2p transport: map the peers' shards (collective)
    }
}

void Circuit::prepareState() {
    allocState();
    checkHq(hq_state_init(deviceStateVec[0], numQubits - MyGlobalVars::bit, MyMPI::rank == 0));
    checkHq(hq_sync());
}

// Everything the reference times as "Time Cost" (circuit.cpp:22-52 there): issue all launches, final sync.
// stateIsGarbage: the state was allocated but not initialised (run()): the first gate group runs as its zero-input variant, which
// neither needs the zero fill nor reads the state; where that variant is unavailable the state is initialised here, inside the
// timed region (the reference's kernelInit memset is outside its "Time Cost", src/circuit.cpp:22-46, so this can only cost us).
int Circuit::execute(std::vector<float>* perGroupMs, bool stateIsGarbage) {
    auto start = chrono::system_clock::now();
    checkHq(hq_timer_start());
    Executor ex(deviceStateVec, numQubits, schedule);
    ex.perGroupMs = perGroupMs;
    if (!stateIsGarbage) ex.run();
    else if (!ex.runFromZero()) {
        checkHq(hq_state_init(deviceStateVec[0], numQubits - MyGlobalVars::bit, MyMPI::rank == 0));
        ex.run();
    }
    auto end = chrono::system_clock::now();
    const int us = (int)chrono::duration_cast<chrono::microseconds>(end - start).count();
    float ms = 0;
    checkHq(hq_timer_stop_ms(&ms));
    // per-group mode re-arms the same event pair for every launch: the whole-run figure is then the wall clock
    lastDeviceMs = perGroupMs ? us * 1e-3 : ms;
    return us;
}

int Circuit::run(bool copy_back, bool destroy) {
    destroyState();
    allocState();   // not initialised: execute(.., stateIsGarbage = true) starts from |0...0> without a zero fill when it can
    const int L = numQubits - MyGlobalVars::bit;
    // HQ_MEASURE_STAGE=1: one Logger line per launch (the reference's -DMEASURE_STAGE, src/executor.cpp:27-30,462-531); every
    // launch is then timed on its own, so nothing overlaps and "Time Cost" is the serialised figure
    std::vector<float> perLaunch;
    const bool measureStage = getenv("HQ_MEASURE_STAGE") != nullptr;
    const int us = execute(measureStage ? &perLaunch : nullptr, /*stateIsGarbage=*/true);
    Logger::add("Time Cost: %d us", us);
    if (measureStage) {
        size_t i = 0;
        int stage = 0;
        for (const auto& lg : schedule.localGroups) {
            const int chunks = 1 << lg.swap.localBit.size();
            for (int c = 0; c < chunks && stage > 0 && !lg.swap.empty(); c++)
                for (const auto& gg : lg.overlapGroups)
                    if (i < perLaunch.size())
                        Logger::add("stage %d chunk %d %s group, %d gates: %.3f ms", stage, c, gg.backend == Backend::BLAS ? "dense" : "tile",
                                    (int)gg.gates.size(), perLaunch[i++]);
            for (const auto& gg : lg.fullGroups)
                if (i < perLaunch.size())
                    Logger::add("stage %d %s group, %d gates: %.3f ms", stage, gg.backend == Backend::BLAS ? "dense" : "tile",
                                (int)gg.gates.size(), perLaunch[i++]);
            stage++;
        }
    }

    collectDump();
    result.clear();
    if (copy_back) {
        const qindex maxAmps = qindex(1) << 29;   // 8 GiB of host memory; larger states stay on the device
        if ((qindex(1) << L) <= maxAmps) {
            result.resize(qindex(1) << L);
            checkHq(hq_state_download(deviceStateVec[0], L, 0, qindex(1) << L, reinterpret_cast<double*>(result.data())));
        } else {
            Logger::add("copy_back skipped: %d local qubits do not fit the host copy budget", L);
        }
    }
    if (destroy) destroyState();
    return us;
}

// The schedule's exchanges alone, back to back, nothing computing: the denominator of "how much of the swap is hidden".
// The data is left wherever the exchanges put it: call it on a state that is not needed any more.
double Circuit::swapAloneMs() {
    if (deviceStateVec.empty() || MyGlobalVars::numGPUs == 1) return 0.0;
    checkHq(hq_sync());
    checkHq(hq_timer_start());
    for (size_t s = 1; s < schedule.localGroups.size(); s++) {
        LocalGroup& lg = schedule.localGroups[s];
        if (lg.swap.empty()) continue;
        hyquas::SwapExec ex(deviceStateVec[0], numQubits - MyGlobalVars::bit, lg.swap, lg.swapPlan);
        ex.begin();
        for (int i = 0; i < (1 << lg.swap.localBit.size()); i++) ex.waitNextChunk();
        ex.end();
    }
    float ms = 0;
    checkHq(hq_timer_stop_ms(&ms));
    return ms;
}

// P(logical qubit reads 0) over the whole distributed state.  A local qubit is a reduction over this shard's "bit clear" half; a
// global qubit selects whole shards (those whose rank bit is 0 contribute their norm).  Collective: every rank calls it.
double Circuit::measure(int logicalQubit) {
    if (deviceStateVec.empty()) return 0.0;
    const int L = numQubits - MyGlobalVars::bit;
    const int p = schedule.finalState.pos.empty() ? logicalQubit : schedule.finalState.pos[logicalQubit];
    double mine = 0;
    if (p < L) checkHq(hq_state_measure(deviceStateVec[0], L, p, &mine))
    else if (!((MyMPI::rank >> (p - L)) & 1)) checkHq(hq_state_norm2(deviceStateVec[0], L, &mine))
    if (MyGlobalVars::numGPUs == 1) return mine;
    std::vector<double> all(MyGlobalVars::numGPUs);
    checkHq(hq_comm_allgather_host(&mine, all.data(), sizeof(double)));
    double total = 0;
    for (double v : all) total += v;
    return total;
}

double Circuit::norm2() {
    double v = 0;
    if (!deviceStateVec.empty()) checkHq(hq_state_norm2(deviceStateVec[0], numQubits - MyGlobalVars::bit, &v));
    return v;
}

size_t Circuit::planBytes() const {
    size_t total = 0;
    for (const auto& lg : schedule.localGroups)
        for (const auto* groups : {&lg.overlapGroups, &lg.fullGroups})
            for (const auto& gg : *groups)
                for (void* p : gg.plans) {
                    int bytes = 0;
                    if (gg.backend == Backend::BLAS) hq_dense_plan_info(static_cast<hq_dense_plan*>(p), nullptr, nullptr, nullptr, nullptr, &bytes);
                    else hq_group_plan_table_bytes(static_cast<hq_group_plan*>(p), &bytes);
                    total += (size_t)std::max(bytes, 0);
                }
    return total;
}

void Circuit::dumpGates() {
    printf("total Gates: %d\n", (int)gates.size());
    for (const Gate& g : gates) {
        for (int i = 0; i < numQubits; i++) {
            if (i == g.controlQubit || i == g.controlQubit2) printf(".  ");
            else if (i == g.targetQubit) printf("%-3s", g.name.c_str());
            else printf("|  ");
        }
        printf("\n");
    }
}

qindex Circuit::toPhysicalID(qindex idx) {
    const auto& pos = schedule.finalState.pos;
    qindex id = 0;
    for (int i = 0; i < numQubits; i++) if (idx >> i & 1) id |= qindex(1) << pos[i];
    return id;
}

qindex Circuit::toLogicID(qindex idx) {
    const auto& pos = schedule.finalState.pos;
    qindex id = 0;
    for (int i = 0; i < numQubits; i++) if (idx >> pos[i] & 1) id |= qindex(1) << i;
    return id;
}

qComplex Circuit::ampAtGPU(qindex idx) {
    const int L = numQubits - MyGlobalVars::bit;
    const qindex id = toPhysicalID(idx);
    qComplex ret = make_qComplex(0.0, 0.0);
    const int owner = (int)(id >> L);
    if (owner == MyMPI::rank) {
        assert(!deviceStateVec.empty());
        checkHq(hq_amp_fetch(deviceStateVec[0], id & ((qindex(1) << L) - 1), reinterpret_cast<double*>(&ret)));
    }
    if (MyGlobalVars::numGPUs > 1) hyquas::bcastAmp(&ret, owner);
    return ret;
}

ResultItem Circuit::ampAt(qindex idx) {
    // With more than one process every rank must take the same path: ampAtGPU ends in a broadcast from the owner (the
    // reference's is a symmetric MPI_Bcast, src/circuit.cpp:95-125).  The host-side shortcuts are single-process only.
    if (MyGlobalVars::numGPUs > 1) return ResultItem(idx, ampAtGPU(idx));
    const qindex id = toPhysicalID(idx);
    if (!result.empty()) return ResultItem(idx, result[id]);
    if (idx < 128 && dumpItems.size() >= 128) return dumpItems[idx];
    return ResultItem(idx, ampAtGPU(idx));
}

// What printState shows (reference circuit.cpp:284-309): logical amplitudes 0..127, then every amplitude with
// |a|^2 > 0.001 and logical index >= 128, ascending.  Captured on the device while the state is alive so that
// nothing of size 2^n ever has to reach the host.
void Circuit::collectDump() {
    const int L = numQubits - MyGlobalVars::bit;
    const qindex localMask = (qindex(1) << L) - 1;
    dumpItems.clear();
    std::vector<ResultItem> mine;
    const int head = (int)std::min<qindex>(128, qindex(1) << numQubits);
    for (int i = 0; i < head; i++) {
        const qindex id = toPhysicalID(i);
        if ((id >> L) != MyMPI::rank) continue;
        qComplex a;
        checkHq(hq_amp_fetch(deviceStateVec[0], id & localMask, reinterpret_cast<double*>(&a)));
        mine.push_back(ResultItem(i, a));
    }
    const int64_t cap = 1024;   // at most 1/0.001 amplitudes can exceed the threshold
    std::vector<int64_t> idx(cap);
    std::vector<double> amp(2 * cap);
    int64_t found = 0;
    checkHq(hq_dump_scan(deviceStateVec[0], L, 0.001, idx.data(), amp.data(), cap, &found));
    assert(found <= cap);
    for (int64_t i = 0; i < found; i++) {
        const qindex logic = toLogicID(idx[i] | (qindex(MyMPI::rank) << L));
        if (logic >= 128) mine.push_back(ResultItem(logic, make_qComplex(amp[2 * i], amp[2 * i + 1])));
    }
    if (MyGlobalVars::numGPUs > 1) hyquas::gatherItems(mine);   // rank 0 receives everybody's items
    std::sort(mine.begin(), mine.end());
    dumpItems.swap(mine);
}

std::string Circuit::stateDump() {
    std::string out;
    for (const auto& item : dumpItems) out += item.str();
    return out;
}

void Circuit::printState() {
    if (MyMPI::rank == 0) {
        fputs(stateDump().c_str(), stdout);
        fflush(stdout);
    }
    // The ranks of a launcher-mode run share one stdout: nobody prints its Logger lines into the middle of rank 0's dump
    // (collective: every rank calls printState, as in the reference's main.cpp:246-249).
    if (MyGlobalVars::numGPUs > 1 && !MyGlobalVars::hostOnly) {
        unsigned char token = 0;
        checkHq(hq_comm_bcast_host(&token, 1, 0));
    }
}

bool Circuit::localShard(double* out) {
    if (deviceStateVec.empty()) return false;
    const int L = numQubits - MyGlobalVars::bit;
    checkHq(hq_state_download(deviceStateVec[0], L, 0, qindex(1) << L, out));
    return true;
}

bool Circuit::fullState(std::vector<qComplex>& out) {
    if (MyGlobalVars::numGPUs != 1 || numQubits > 30) return false;
    const qindex N = qindex(1) << numQubits;
    std::vector<qComplex> phys;
    if (!result.empty()) phys = result;
    else if (!deviceStateVec.empty()) {
        phys.resize(N);
        checkHq(hq_state_download(deviceStateVec[0], numQubits, 0, N, reinterpret_cast<double*>(phys.data())));
    } else return false;
    bool identity = true;
    for (int i = 0; i < numQubits; i++) identity &= schedule.finalState.pos[i] == i;
    if (identity) { out.swap(phys); return true; }
    out.resize(N);
    for (qindex i = 0; i < N; i++) out[i] = phys[toPhysicalID(i)];
    return true;
}
