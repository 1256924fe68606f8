// Schedule data model: what the partitioner hands to the executor.
//   Schedule = sequence of LocalGroups (communication stages); a LocalGroup = qubit layout for the stage +
//   the global<->local swap that establishes it + gate groups; a GateGroup = gates applied by ONE kernel
//   launch (one sweep over the local state) together with the device plan for this process' GPU.
// Same vocabulary as the reference (src/schedule.h:9-104).  Differences that matter:
//   * every kernel is in place and never permutes amplitudes inside a GPU, so a GateGroup does not change
//     the layout (the reference's BLAS groups do, schedule.cpp:339-382) and there are no cuTT plans;
//   * the layout only changes at stage boundaries, by explicit SwapSteps (hyquas::SwapPlan).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "gate.h"
#include "utils.h"

enum class Backend { None, PerGate, BLAS };   // PerGate = tile (OShareMem-class) kernel, BLAS = fused dense (TransMM) kernel
std::string to_string(Backend b);

struct State {
    std::vector<int> pos;     // pos[logical qubit] = physical bit of the amplitude index
    std::vector<int> layout;  // layout[physical bit] = logical qubit
    State() = default;
    explicit State(int numQubits) {
        for (int i = 0; i < numQubits; i++) { pos.push_back(i); layout.push_back(i); }
    }
    void swapPhysical(int a, int b) {   // exchange the logical qubits sitting at physical bits a and b
        std::swap(layout[a], layout[b]);
        pos[layout[a]] = a;
        pos[layout[b]] = b;
    }
};

namespace hyquas {
// One global<->local exchange: physical local bit localBit[i] trades places with global bit globalBit[i]
// (globalBit counted from 0 = physical bit L).  Before the exchange the outgoing qubits are brought to
// localBit[] by in-place bit swaps (localPerm: list of physical (a, b) pairs applied in order).
struct SwapPlan {
    std::vector<std::pair<int, int>> localPerm;
    std::vector<int> localBit;
    std::vector<int> globalBit;
    bool empty() const { return localBit.empty(); }
};
}  // namespace hyquas

// One dense matrix of a BLAS-backend group: the gates it fuses and the logical qubits it spans.
struct DenseBlock {
    qindex qubits = 0;
    std::vector<Gate> gates;
};

struct GateGroup {
    std::vector<Gate> gates;
    qindex relatedQubits = 0;     // logical qubits the kernel keeps in its tile / matrix
    Backend backend = Backend::PerGate;
    State state;                  // layout while (and after) this group runs
    int matQubit = 0;             // BLAS: matrix qubits of the largest block
    std::vector<DenseBlock> blocks;   // BLAS: the dense matrices applied, in order, by this one launch
    double predictedMs = 0;       // evaluator's estimate

    // device side, filled by Circuit::compile() for this process
    uint64_t tileMask = 0;                  // physical bits of the tile
    std::vector<void*> plans;               // full group: 1 plan; overlap (per-chunk) group: one per chunk
    bool contains(int i) const { return (relatedQubits >> i) & 1; }
};

struct LocalGroup {
    State state;                            // layout during this stage
    hyquas::SwapPlan swap;                  // how the stage's layout is reached from the previous one (empty for stage 0)
    std::vector<GateGroup> overlapGroups;   // run per received chunk, overlapped with the exchange
    std::vector<GateGroup> fullGroups;
    qindex relatedQubits = 0;               // logical qubits that are local in this stage
    void* swapPlan = nullptr;               // device-side hq_swap_plan of this process (built by Executor::prepare)
    bool contains(int i) const { return (relatedQubits >> i) & 1; }
};

struct Schedule {
    std::vector<LocalGroup> localGroups;
    State finalState;
    void dump(int numQubits) const;
    int numGroups() const;
};
