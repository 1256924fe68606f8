#include "swap.h"

#include <unistd.h>

#include <chrono>
#include <cstring>
#include <fstream>
#include <thread>

#include "circuit.h"

namespace hyquas {

// Rendezvous for the NCCL unique id when no embedding host provides one: rank 0 writes it to
// /tmp/hyquas_b200_nccl_<launcher pid>_<MASTER_PORT>, the other ranks (children of the same launcher) poll for it.
void commInitFromEnv() {
    int world = 1, rank = 0;
    checkHq(hq_comm_info(&world, &rank));
    if (world == MyGlobalVars::numGPUs && world > 1) return;   // already initialised through hq_comm_init
    const char* port = getenv("MASTER_PORT");
    char path[256];
    snprintf(path, sizeof(path), "/tmp/hyquas_b200_nccl_%d_%s", (int)getppid(), port ? port : "0");
    unsigned char id[128];
    if (MyMPI::rank == 0) {
        checkHq(hq_comm_unique_id(id));
        std::string tmp = std::string(path) + ".tmp";
        std::ofstream(tmp, std::ios::binary).write(reinterpret_cast<const char*>(id), sizeof(id));
        rename(tmp.c_str(), path);
    } else {
        for (int tries = 0;; tries++) {
            std::ifstream in(path, std::ios::binary);
            if (in && in.read(reinterpret_cast<char*>(id), sizeof(id))) break;
            if (tries > 6000) {
                fprintf(stderr, "timed out waiting for the NCCL id file %s\n", path);
                exit(1);
            }
            std::this_thread::sleep_for(std::chrono::milliseconds(10));
        }
    }
    checkHq(hq_comm_init(MyGlobalVars::numGPUs, MyMPI::rank, id));
    if (MyMPI::rank == 0) {   // everybody has joined once the communicator exists
        unsigned char token = 0;
        checkHq(hq_comm_bcast_host(&token, 1, 0));
        unlink(path);
    } else {
        unsigned char token = 0;
        checkHq(hq_comm_bcast_host(&token, 1, 0));
    }
}

void bcastAmp(qComplex* amp, int ownerRank) { checkHq(hq_comm_bcast_host(amp, sizeof(qComplex), ownerRank)); }

void gatherItems(std::vector<ResultItem>& items) {
    // sizes first, then fixed-size slots: the dump holds at most 128 + 1000 items in total
    const int world = MyGlobalVars::numGPUs;
    long long mine = (long long)items.size();
    std::vector<long long> counts(world);
    checkHq(hq_comm_allgather_host(&mine, counts.data(), sizeof(long long)));
    long long cap = 1;
    for (long long c : counts) cap = std::max(cap, c);
    struct Slot { long long idx; double re, im; };
    std::vector<Slot> send(cap), recv((size_t)cap * world);
    for (size_t i = 0; i < items.size(); i++) send[i] = {items[i].idx, items[i].amp.x, items[i].amp.y};
    checkHq(hq_comm_allgather_host(send.data(), recv.data(), sizeof(Slot) * cap));
    std::vector<ResultItem> all;
    if (MyMPI::rank == 0)
        for (int r = 0; r < world; r++)
            for (long long i = 0; i < counts[r]; i++) {
                const Slot& s = recv[(size_t)r * cap + i];
                all.push_back(ResultItem(s.idx, make_qComplex(s.re, s.im)));
            }
    items.swap(all);
}

SwapExec::SwapExec(qComplex* s, int L, const SwapPlan& p, void* dp) : state(s), numLocal(L), plan(p), devicePlan(dp) {}

void SwapExec::begin() {
    if (!plan.localPerm.empty()) {
        std::vector<int> a, b;
        for (auto& pr : plan.localPerm) { a.push_back(pr.first); b.push_back(pr.second); }
        checkHq(hq_state_bitswap(state, numLocal, (int)a.size(), a.data(), b.data()));
    }
    checkHq(hq_swap_begin(static_cast<hq_swap_plan*>(devicePlan), state));
}

int SwapExec::waitNextChunk() {
    int chunk = 0;
    checkHq(hq_swap_wait_chunk(static_cast<hq_swap_plan*>(devicePlan), &chunk));
    return chunk;
}

void SwapExec::end() { checkHq(hq_swap_end(static_cast<hq_swap_plan*>(devicePlan))); }

}  // namespace hyquas
