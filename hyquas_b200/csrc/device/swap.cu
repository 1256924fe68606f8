// Multi-GPU layer of the device library: NCCL communicator + the in-place, chunked global<->local qubit swap.
//
// What it replaces: Executor::transpose + Executor::all2all + sliceBarrier (src/executor.cpp:59-179,650-659) and the
// NCCL bootstrap in MyGlobalVars::init (src/utils.cpp:46-58).  The reference transposes the whole local state into a
// SECOND full-size buffer (cuTT) and then copies/sends 2^g contiguous parts; here
//   * process model = one process per GPU, one NCCL communicator over NVLink 5 / NVSwitch;
//   * the swap exchanges the top k local bits with k global bits: the local state is 2^k contiguous CHUNKS of
//     2^(L-k) amplitudes; chunk c of rank r trades places with chunk c(r) of rank r(c)  (a pairwise exchange, so it is
//     done in place, step xr = 1 .. 2^k-1 pairing every rank with rank^xr: a perfect matching per step, like the
//     xr-major slice order of the reference, src/executor.cpp:66-179);
//   * a chunk moves in PIECES (default 64 MiB) through a two-slot staging ring: ncclSend(piece) + ncclRecv(slot) in one
//     group on the comm stream, then an un-stage copy on a third stream while the next piece is in flight, so the
//     extra memory is 2 pieces instead of a second state vector;
//   * one CUDA event per chunk: the compute stream waits for exactly the chunk it is about to run the overlap
//     groups on, while later chunks are still on the wire.
// NCCL is bound with dlopen (libnccl.so.2) so that single-GPU use never needs it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <vector>

#include "hq_internal.h"

namespace hq {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

struct Comm {
    NcclApi api;
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    cudaStream_t unstage = nullptr;     // staging slot -> state copies
    double2* staging = nullptr;         // 2 slots of piece_amps
    uint64_t piece_amps = 0;
    cudaEvent_t slot_filled[2] = {nullptr, nullptr}, slot_free[2] = {nullptr, nullptr};
};
static Comm& cm() {
    static Comm c;
    return c;
}

static int nccl_fail(ncclResult_t r, const char* what, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s in file %s, line %d: %04d %s", what, __FILE__, line, (int)r,
             cm().api.GetErrorString ? cm().api.GetErrorString(r) : "nccl error");
    set_error(buf);
    return HQ_ERR_NCCL;
}
#define HQ_NCCL(stmt)                                                      \
    do {                                                                   \
        ncclResult_t _r = (stmt);                                          \
        if (_r != ncclSuccess) return ::hq::nccl_fail(_r, #stmt, __LINE__); \
    } while (0)

static int load_nccl() {
    NcclApi& a = cm().api;
    if (a.handle) return HQ_OK;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (a.handle) break;
    }
    if (!a.handle) {
        set_error(std::string("cannot load NCCL: ") + dlerror());
        return HQ_ERR_NCCL;
    }
    bool ok = true;
    auto sym = [&](const char* n) {
        void* p = dlsym(a.handle, n);
        ok &= p != nullptr;
        return p;
    };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
    a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
    a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(sym("ncclBroadcast"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    if (!ok) {
        set_error("libnccl lacks a required symbol");
        return HQ_ERR_NCCL;
    }
    return HQ_OK;
}

// In-place permutation of the local state by a product of DISJOINT physical-bit transpositions (a_i <-> b_i): the
// index map is an involution, so element x trades places with pi(x) and only the pair's smaller index does the work.
struct BitSwaps {
    int n;
    uint8_t a[8], b[8];
};
__global__ void bitswap_kernel(double2* s, uint64_t n, const BitSwaps bs) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += stride) {
        uint64_t y = x;
#pragma unroll 1
        for (int i = 0; i < bs.n; ++i) {
            const uint64_t d = ((x >> bs.a[i]) ^ (x >> bs.b[i])) & 1ull;
            y ^= (d << bs.a[i]) | (d << bs.b[i]);
        }
        if (x < y) {
            const double2 u = s[x], v = s[y];
            s[x] = v;
            s[y] = u;
        }
    }
}

__global__ void unstage_kernel(const double2* __restrict__ src, double2* __restrict__ dst, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

}  // namespace hq

using namespace hq;

struct hq_swap_plan {
    int L = 0, k = 0;
    int myc = 0;                      // my value of the swapped global bits = the chunk that stays
    std::vector<int> peer;            // peer[xr] = rank exchanged with at step xr (xr = 1 .. 2^k-1)
    std::vector<cudaEvent_t> landed;  // landed[c]: chunk c holds its post-swap contents
    cudaEvent_t compute_done = nullptr, all_done = nullptr;
    int next = 0;
};

extern "C" int hq_comm_unique_id(unsigned char out[128]) {
    HQ_REQUIRE(out != nullptr, "null out pointer");
    int rc = load_nccl();
    if (rc != HQ_OK) return rc;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    HQ_NCCL(cm().api.GetUniqueId(&id));
    std::memcpy(out, &id, 128);
    return HQ_OK;
}

extern "C" int hq_comm_init(int world, int rank, const unsigned char id_bytes[128]) {
    HQ_REQUIRE(rt().ready, "hq_init() must be called before hq_comm_init()");
    HQ_REQUIRE(world >= 1 && rank >= 0 && rank < world && id_bytes, "bad arguments to hq_comm_init");
    Comm& c = cm();
    if (c.comm) return HQ_OK;
    int rc = load_nccl();
    if (rc != HQ_OK) return rc;
    ncclUniqueId id;
    std::memcpy(&id, id_bytes, 128);
    HQ_NCCL(c.api.CommInitRank(&c.comm, world, id, rank));
    c.world = world;
    c.rank = rank;
    HQ_CUDA(cudaStreamCreateWithFlags(&c.unstage, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        HQ_CUDA(cudaEventCreateWithFlags(&c.slot_filled[i], cudaEventDisableTiming));
        HQ_CUDA(cudaEventCreateWithFlags(&c.slot_free[i], cudaEventDisableTiming));
    }
    return HQ_OK;
}

extern "C" int hq_comm_info(int* world, int* rank) {
    if (world) *world = cm().world;
    if (rank) *rank = cm().rank;
    return HQ_OK;
}

extern "C" int hq_comm_destroy(void) {
    Comm& c = cm();
    if (!c.comm) return HQ_OK;
    cudaStreamSynchronize(rt().comm);
    cudaStreamSynchronize(c.unstage);
    c.api.CommDestroy(c.comm);
    c.comm = nullptr;
    if (c.staging) cudaFree(c.staging);
    c.staging = nullptr;
    for (int i = 0; i < 2; ++i) {
        cudaEventDestroy(c.slot_filled[i]);
        cudaEventDestroy(c.slot_free[i]);
    }
    cudaStreamDestroy(c.unstage);
    c.world = 1;
    c.rank = 0;
    return HQ_OK;
}

// Small host-buffer collectives for the control plane of printState (amplitude of one index, dump items).
extern "C" int hq_comm_bcast_host(void* buf, size_t bytes, int root) {
    Comm& c = cm();
    HQ_REQUIRE(c.comm && buf, "communicator not initialised");
    void* d = nullptr;
    HQ_CUDA(cudaMalloc(&d, bytes));
    HQ_CUDA(cudaMemcpyAsync(d, buf, bytes, cudaMemcpyHostToDevice, rt().comm));
    HQ_NCCL(c.api.Broadcast(d, d, bytes, ncclUint8, root, c.comm, rt().comm));
    HQ_CUDA(cudaMemcpyAsync(buf, d, bytes, cudaMemcpyDeviceToHost, rt().comm));
    HQ_CUDA(cudaStreamSynchronize(rt().comm));
    cudaFree(d);
    return HQ_OK;
}

extern "C" int hq_comm_allgather_host(const void* send, void* recv, size_t bytes_per_rank) {
    Comm& c = cm();
    HQ_REQUIRE(c.comm && send && recv, "communicator not initialised");
    void *ds = nullptr, *dr = nullptr;
    HQ_CUDA(cudaMalloc(&ds, bytes_per_rank));
    HQ_CUDA(cudaMalloc(&dr, bytes_per_rank * c.world));
    HQ_CUDA(cudaMemcpyAsync(ds, send, bytes_per_rank, cudaMemcpyHostToDevice, rt().comm));
    HQ_NCCL(c.api.AllGather(ds, dr, bytes_per_rank, ncclUint8, c.comm, rt().comm));
    HQ_CUDA(cudaMemcpyAsync(recv, dr, bytes_per_rank * c.world, cudaMemcpyDeviceToHost, rt().comm));
    HQ_CUDA(cudaStreamSynchronize(rt().comm));
    cudaFree(ds);
    cudaFree(dr);
    return HQ_OK;
}

extern "C" int hq_state_bitswap(void* state, int L, int npairs, const int* a, const int* b) {
    HQ_REQUIRE(rt().ready && state && npairs >= 0 && npairs <= 8, "bad arguments to hq_state_bitswap");
    if (npairs == 0) return HQ_OK;
    BitSwaps bs{};
    uint64_t used = 0;
    for (int i = 0; i < npairs; ++i) {
        HQ_REQUIRE(a[i] >= 0 && a[i] < L && b[i] >= 0 && b[i] < L && a[i] != b[i], "bit swap outside the local state");
        HQ_REQUIRE(!(used >> a[i] & 1) && !(used >> b[i] & 1), "bit swaps must be disjoint");
        used |= (1ull << a[i]) | (1ull << b[i]);
        bs.a[i] = (uint8_t)a[i];
        bs.b[i] = (uint8_t)b[i];
    }
    bs.n = npairs;
    const uint64_t n = 1ull << L;
    const int block = 256;
    const int grid = (int)std::min<uint64_t>((n + block - 1) / block, (uint64_t)rt().sm_count * 16);
    bitswap_kernel<<<grid, block, 0, rt().compute>>>(static_cast<double2*>(state), n, bs);
    HQ_CUDA(cudaGetLastError());
    return HQ_OK;
}

// local_bits must be the top k local positions (L-k .. L-1, ascending): chunks are then contiguous.
// global_bits[i] (0-based above L) is the global bit traded with local_bits[i].
extern "C" int hq_swap_plan_create(int L, int k, const int* local_bits, const int* global_bits, hq_swap_plan** out) {
    Comm& c = cm();
    HQ_REQUIRE(out && k >= 1 && k <= 6 && L > k, "bad arguments to hq_swap_plan_create");
    HQ_REQUIRE(c.comm != nullptr, "hq_comm_init() has not been called");
    int g = 0;
    while ((1 << g) < c.world) ++g;
    HQ_REQUIRE(k <= g, "cannot swap more bits than there are global qubits");
    auto* p = new hq_swap_plan();
    p->L = L;
    p->k = k;
    for (int i = 0; i < k; ++i) {
        if (local_bits[i] != L - k + i || global_bits[i] < 0 || global_bits[i] >= g) {
            delete p;
            set_error("swap plan: local bits must be the top k local positions and global bits must exist");
            return HQ_ERR_ARG;
        }
        p->myc |= ((c.rank >> global_bits[i]) & 1) << i;
    }
    p->peer.assign(1 << k, c.rank);
    for (int xr = 1; xr < (1 << k); ++xr) {
        const int ch = p->myc ^ xr;
        int r = c.rank;
        for (int i = 0; i < k; ++i) r = (r & ~(1 << global_bits[i])) | (((ch >> i) & 1) << global_bits[i]);
        p->peer[xr] = r;
    }
    p->landed.resize(1 << k);
    for (auto& e : p->landed) HQ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    HQ_CUDA(cudaEventCreateWithFlags(&p->compute_done, cudaEventDisableTiming));
    HQ_CUDA(cudaEventCreateWithFlags(&p->all_done, cudaEventDisableTiming));
    *out = p;
    return HQ_OK;
}

extern "C" int hq_swap_plan_destroy(hq_swap_plan* p) {
    if (!p) return HQ_OK;
    for (auto& e : p->landed) cudaEventDestroy(e);
    cudaEventDestroy(p->compute_done);
    cudaEventDestroy(p->all_done);
    delete p;
    return HQ_OK;
}

// Enqueue the whole exchange on the comm stream (after everything already queued on the compute stream).
extern "C" int hq_swap_begin(hq_swap_plan* p, void* state_v) {
    Comm& c = cm();
    HQ_REQUIRE(p && state_v && c.comm, "bad arguments to hq_swap_begin");
    double2* state = static_cast<double2*>(state_v);
    const uint64_t chunk_amps = 1ull << (p->L - p->k);
    uint64_t piece = 1ull << 22;   // 64 MiB of amplitudes
    if (const char* e = getenv("HQ_SWAP_PIECE_LOG2")) piece = 1ull << std::max(10, std::min(atoi(e), 30));
    piece = std::min(piece, chunk_amps);
    if (c.piece_amps != piece) {
        if (c.staging) {
            HQ_CUDA(cudaStreamSynchronize(c.unstage));
            HQ_CUDA(cudaFree(c.staging));
        }
        HQ_CUDA(cudaMalloc(&c.staging, 2 * piece * sizeof(double2)));
        c.piece_amps = piece;
        for (int s = 0; s < 2; ++s) HQ_CUDA(cudaEventRecord(c.slot_free[s], c.unstage));
    }
    cudaStream_t comm = rt().comm;
    HQ_CUDA(cudaEventRecord(p->compute_done, rt().compute));
    HQ_CUDA(cudaStreamWaitEvent(comm, p->compute_done, 0));
    HQ_CUDA(cudaEventRecord(p->landed[p->myc], comm));   // the chunk that stays is ready as soon as compute is
    const uint64_t npieces = chunk_amps / piece;
    uint64_t seq = 0;
    for (int xr = 1; xr < (1 << p->k); ++xr) {
        const int ch = p->myc ^ xr, peer = p->peer[xr];
        double2* chunk = state + (uint64_t)ch * chunk_amps;
        for (uint64_t q = 0; q < npieces; ++q, ++seq) {
            const int slot = (int)(seq & 1);
            double2* stage = c.staging + (uint64_t)slot * piece;
            HQ_CUDA(cudaStreamWaitEvent(comm, c.slot_free[slot], 0));
            HQ_NCCL(c.api.GroupStart());
            HQ_NCCL(c.api.Send(chunk + q * piece, piece * 2, ncclDouble, peer, c.comm, comm));
            HQ_NCCL(c.api.Recv(stage, piece * 2, ncclDouble, peer, c.comm, comm));
            HQ_NCCL(c.api.GroupEnd());
            HQ_CUDA(cudaEventRecord(c.slot_filled[slot], comm));
            HQ_CUDA(cudaStreamWaitEvent(c.unstage, c.slot_filled[slot], 0));
            unstage_kernel<<<rt().sm_count * 2, 512, 0, c.unstage>>>(stage, chunk + q * piece, piece);
            HQ_CUDA(cudaGetLastError());
            HQ_CUDA(cudaEventRecord(c.slot_free[slot], c.unstage));
        }
        HQ_CUDA(cudaEventRecord(p->landed[ch], c.unstage));
    }
    HQ_CUDA(cudaEventRecord(p->all_done, c.unstage));
    p->next = 0;
    return HQ_OK;
}

// Makes the compute stream wait for the next chunk in arrival order; returns its index.
extern "C" int hq_swap_wait_chunk(hq_swap_plan* p, int* chunk) {
    HQ_REQUIRE(p && chunk && p->next < (1 << p->k), "no chunk left to wait for");
    const int ch = p->myc ^ p->next;
    HQ_CUDA(cudaStreamWaitEvent(rt().compute, p->landed[ch], 0));
    *chunk = ch;
    p->next++;
    return HQ_OK;
}

// After this the compute stream is ordered behind the whole exchange (and the next swap behind compute).
extern "C" int hq_swap_end(hq_swap_plan* p) {
    HQ_REQUIRE(p != nullptr, "null swap plan");
    HQ_CUDA(cudaStreamWaitEvent(rt().compute, p->all_done, 0));
    return HQ_OK;
}
