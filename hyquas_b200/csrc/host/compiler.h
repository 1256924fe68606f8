// Hybrid partitioner: gate list -> Schedule.
//   1. split the circuit into communication stages (each stage's non-diagonal targets fit in the local qubits),
//   2. plan the global<->local swap between consecutive stages,
//   3. cut each stage into gate groups (one kernel launch = one sweep each) and pick the backend per group
//      from the evaluator's measured cost model.
// Plays the role of the reference's Compiler / SimpleCompiler / AdvanceCompiler (src/compiler.h:9-61,
// src/compiler.cpp:70-452); the algorithms are new (frontier scans with commutation-aware blocking and a
// greedy qubit-gain search instead of the std::bitset reachability DP).
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

#include "gate.h"
#include "schedule.h"
#include "utils.h"

class Compiler {
public:
    Compiler(int numQubits, std::vector<Gate> inputGates);
    Schedule run();

    // knobs (defaults come from the device layer / environment)
    int tileBits;      // qubits per tile of the gate-group kernel
    int pinnedBits;    // lowest physical bits always kept in the tile (contiguous-run length of HBM accesses)
    int maxGroupGates; // cap on gates per group
    bool rebalanceGroups; // adjacent tile groups trade gates to even out compute-bound and sweep-bound launches (HQ_REBALANCE)
    int cutVariants;   // most differently seeded greedy cuts tried per stage; the cheapest predicted total wins (HQ_CUT_VARIANTS, 1 = off)
    bool cutBothWays;  // try the greedy cut from both ends of a stage, keep the cheaper predicted total
    bool enableOverlap; // per-chunk groups under the exchange (reference: ENABLE_OVERLAP)
    int backendMode;   // 1 = tile kernel only, 3 = dense kernel only, 4 = hybrid (reference: -DBACKEND=group|blas|mix)
    int matLimit;      // largest dense block in qubits (reference: -DMAT, BLAS_MAT_LIMIT)
    double overlapSlack; // deferred work may take up to this multiple of the predicted exchange time
private:
    struct Stage { std::vector<Gate> gates; qindex locals; };
    std::vector<Stage> splitStages() const;
    std::vector<Stage> splitStagesVariant(int variant) const;
    GateGroup denseCandidate(const std::vector<Gate>& gates, const std::vector<int>& remaining, const State& state, int numLocal, qindex exclude) const;
    std::vector<GateGroup> cutGroups(const std::vector<Gate>& gates, const State& state, int numLocal, qindex exclude, bool search = true) const;
    std::vector<GateGroup> cutGroupsBothWays(const std::vector<Gate>& gates, const State& state, int numLocal, qindex exclude) const;
    void rebalance(std::vector<GateGroup>& groups, int nEff) const;
    void absorbCrumbs(std::vector<GateGroup>& groups, int nEff) const;
    std::vector<GateGroup> cutGroupsGreedy(const std::vector<Gate>& gates, const State& state, int numLocal, qindex exclude, int variant = 0,
                                           bool tileOnly = false) const;
    int trialsFor(size_t numGates) const;
    unsigned long long cutKey(const std::vector<Gate>& gates, const State& state, int numLocal, qindex exclude) const;
    unsigned long long circuitKey() const;
    void loadWisdom();
    void saveWisdom() const;
    // winning seed of every searched cut of this circuit (0 = the plain greedy cut), kept across compiles: see cutGroups
    mutable std::unordered_map<unsigned long long, int> wisdom;
    mutable bool wisdomDirty = false;
    int numQubits;
    int numLocal;
    std::vector<Gate> gates;
};

namespace hyquas {
// Frontier scan shared by the stage splitter, the group cutter and (on the device side) the round builder:
// which of `gates` (indices into it, in program order) can run now if exactly the qubits in `tileSet` may be
// non-diagonal targets?  Gates that cannot run block later gates they do not commute with.
std::vector<int> runnableGates(const std::vector<Gate>& gates, const std::vector<int>& order, qindex tileSet, int cap);
std::vector<int> runnableDense(const std::vector<Gate>& gates, const std::vector<int>& order, qindex qset, int cap);
// "" when `schedule` is a valid way to run `gates` (every gate once, non-commuting pairs in order, targets local / in tile / off
// the exchanged positions), else what is wrong with it
std::string checkSchedule(const Schedule& schedule, const std::vector<Gate>& gates, int numQubits, int numLocal, int tileBits);
hyquas::SwapPlan planSwap(State& state, qindex newLocals, int numQubits, int numLocal, bool anyBit);
}
