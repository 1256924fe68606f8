#include "swap.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <ctime>
#include <cstring>
#include <fstream>
#include <thread>

#include "circuit.h"

namespace hyquas {

// Rendezvous for the NCCL unique id when no embedding host provides one.  Rank 0 writes it to a file named after the launcher's
// pid, MASTER_PORT and the launcher's run id, created exclusively (O_EXCL | O_NOFOLLOW, mode 0600) after removing any stale
// file of an earlier run; the other ranks (children of the same launcher) poll for a file that belongs to this user and is
// recent, so a leftover of a crashed run with the same pid / port is never taken for this run's id.
static std::string rendezvousPath() {
    const char* port = getenv("MASTER_PORT");
    const char* run = getenv("TORCHELASTIC_RUN_ID");
    const char* dir = getenv("XDG_RUNTIME_DIR");
    std::string d = (dir && access(dir, W_OK) == 0) ? dir : "/tmp";
    return d + "/hyquas_b200_nccl_" + std::to_string((int)getuid()) + "_" + std::to_string((int)getppid()) + "_" + (port ? port : "0") +
           "_" + (run ? run : "norun");
}

void commInitFromEnv() {
    int world = 1, rank = 0;
    checkHq(hq_comm_info(&world, &rank));
    if (world == MyGlobalVars::numGPUs && world > 1) return;   // already initialised through hq_comm_init
    const std::string path = rendezvousPath();
    unsigned char id[128];
    if (MyMPI::rank == 0) {
        checkHq(hq_comm_unique_id(id));
        unlink(path.c_str());
        const std::string tmp = path + ".tmp";
        unlink(tmp.c_str());
        const int fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL | O_NOFOLLOW, 0600);
        if (fd < 0 || write(fd, id, sizeof(id)) != (ssize_t)sizeof(id)) {
            fprintf(stderr, "cannot write the NCCL id file %s\n", tmp.c_str());
            exit(1);
        }
        close(fd);
        rename(tmp.c_str(), path.c_str());
    } else {
        const time_t started = time(nullptr);
        for (int tries = 0;; tries++) {
            struct stat st;
            const int fd = open(path.c_str(), O_RDONLY | O_NOFOLLOW);
            bool ok = false;
            if (fd >= 0) {
                // ours, a regular file, and written no earlier than a minute before this process began to wait
                ok = fstat(fd, &st) == 0 && st.st_uid == getuid() && S_ISREG(st.st_mode) && st.st_mtime + 60 >= started &&
                     read(fd, id, sizeof(id)) == (ssize_t)sizeof(id);
                close(fd);
            }
            if (ok) break;
            if (tries > 6000) {
                fprintf(stderr, "timed out waiting for the NCCL id file %s\n", path.c_str());
                exit(1);
            }
            std::this_thread::sleep_for(std::chrono::milliseconds(10));
        }
    }
    checkHq(hq_comm_init(MyGlobalVars::numGPUs, MyMPI::rank, id));
    unsigned char token = 0;   // everybody has joined once the communicator exists
    checkHq(hq_comm_bcast_host(&token, 1, 0));
    if (MyMPI::rank == 0) unlink(path.c_str());
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
