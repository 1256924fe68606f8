// Multi-GPU plumbing of the host layer: communicator bootstrap and the global<->local qubit swap
// (the reference's Executor::transpose + all2all + sliceBarrier, src/executor.cpp:59-179,650-659).
#pragma once
#include <vector>
#include "schedule.h"

struct ResultItem;
namespace hyquas {

void commInitFromEnv();
void bcastAmp(qComplex* amp, int ownerRank);            // rank `ownerRank` -> everybody
void gatherItems(std::vector<ResultItem>& items);  // everybody -> rank 0 (others end up empty)   // WORLD_SIZE/RANK + file rendezvous for the NCCL unique id

// Executes one SwapPlan in place on this process' shard, chunk by chunk, on the comm stream.
class SwapExec {
public:
    SwapExec(qComplex* state, int numLocal, const SwapPlan& plan);
    void begin();              // local bit swaps (compute stream) + enqueue the chunked exchange (comm stream)
    int waitNextChunk();       // makes the compute stream wait for the next landed chunk; returns its index
    void end();
private:
    qComplex* state;
    int numLocal;
    const SwapPlan& plan;
    int next = 0;
};

}  // namespace hyquas
