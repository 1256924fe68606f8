// Multi-GPU plumbing of the host layer: communicator bootstrap and the global<->local qubit swap
// (the reference's Executor::transpose + all2all + sliceBarrier, src/executor.cpp:59-179,650-659).
#pragma once
#include <vector>
#include "schedule.h"

struct ResultItem;
namespace hyquas {

// Creates the NCCL communicator from WORLD_SIZE / RANK unless the embedding host (bench.py through
// hq_comm_init) already did.  The unique id travels through a file named after the launcher's pid and MASTER_PORT.
void commInitFromEnv();
void bcastAmp(qComplex* amp, int ownerRank);       // rank `ownerRank` -> everybody
void gatherItems(std::vector<ResultItem>& items);  // everybody -> rank 0 (the others end up empty)

// Executes one SwapPlan in place on this process' shard, chunk by chunk.
class SwapExec {
public:
    SwapExec(qComplex* state, int numLocal, const SwapPlan& plan, void* devicePlan);
    void begin();              // local bit swaps (compute stream) + enqueue the chunked exchange (comm stream)
    int waitNextChunk();       // makes the compute stream wait for the next landed chunk; returns its index
    void end();
private:
    qComplex* state;
    int numLocal;
    const SwapPlan& plan;
    void* devicePlan;
};

}  // namespace hyquas
