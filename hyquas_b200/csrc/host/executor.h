// Schedule interpreter for one process / one GPU.
//   prepare(): lowers every gate group for THIS shard of the state (global controls resolved, diagonal gates on
//              global qubits folded into phases) and builds the device plans -- done once, at compile time;
//   run():     issues the launches: [swap + per-chunk overlap groups] then full groups, stage by stage.
// Role of the reference's Executor (src/executor.h:10-54, src/executor.cpp:33-57,190-403,462-575) minus
// everything its kernels needed from the host per launch (threadBias tables, constant-memory uploads,
// cuTT/cuBLAS calls).
#pragma once
#include <vector>

#include "schedule.h"
#include "utils.h"

class Executor {
public:
    Executor(std::vector<qComplex*> deviceStateVec, int numQubits, Schedule& schedule);
    void run();
    // The state holds garbage and must become (schedule applied to |0...0>): the first gate group runs as its zero-input variant,
    // which reads nothing; false when that variant is unavailable (the caller zero-fills and calls run()).
    bool runFromZero();
    static bool zeroInputEnabled();
    std::vector<float>* perGroupMs = nullptr;   // when set: one CUDA-event timing per gate-group launch (MEASURE_STAGE)
    static void prepare(Schedule& schedule, int numQubits, bool hostOnly = false);   // build device plans (idempotent)
    static void release(Schedule& schedule);                  // destroy device plans
    // Lower one logical gate for the sub-state whose physical index bits in `fixedMask` equal `fixedValue` (rank bits,
    // plus the swapped positions for a per-chunk launch).  Returns false when the gate acts as identity there.
    static bool lowerGate(const Gate& gate, const State& state, qindex fixedMask, qindex fixedValue, hq_gate& out);
private:
    void applyGateGroup(GateGroup& gg, int chunk);
    bool firstFromZero = false;
    void finalize();
    std::vector<qComplex*> deviceStateVec;
    int numQubits;
    Schedule& schedule;
};
