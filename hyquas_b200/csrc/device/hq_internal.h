// Internal declarations shared by the device-layer translation units (not installed).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

#include "hyquas_b200.h"

namespace hq {

// ---- error plumbing ----------------------------------------------------------------------------
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define HQ_CUDA(stmt)                                                        \
    do {                                                                     \
        cudaError_t _e = (stmt);                                             \
        if (_e != cudaSuccess) return ::hq::cuda_fail(_e, #stmt, __FILE__, __LINE__); \
    } while (0)

#define HQ_REQUIRE(cond, msg)                       \
    do {                                            \
        if (!(cond)) {                              \
            ::hq::set_error(std::string(msg));      \
            return HQ_ERR_ARG;                      \
        }                                           \
    } while (0)

// ---- per-process runtime state (one GPU per process) -------------------------------------------
struct Runtime {
    bool ready = false;
    int device = -1;
    int sm_count = 0;
    cudaStream_t compute = nullptr;   // gate groups
    cudaStream_t comm = nullptr;      // global<->local qubit swaps
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    int tile_bits = 12;               // K of the gate-group kernel
    int reserved_ctas = 0;            // CTAs of a swap kernel in flight on the comm stream (one per SM)
    bool relaxed_regs = false;        // trade resident CTAs for registers in the gate-group kernel
    // the last state vector freed, kept for the next hq_state_alloc of the same size: cudaMalloc + cudaFree of 16 GiB cost
    // ~35 ms per Circuit::run (bench e2e breakdown), more than two sweeps of the state.  HQ_STATE_CACHE=0 turns it off.
    void* cached_state = nullptr;
    size_t cached_bytes = 0;
    bool state_cache = true;
};
Runtime& rt();

// Small-buffer device allocator for plan tables and scratch (sizes rounded up to a power of two >= 4 KiB; freed buffers are
// kept on per-size free lists and only go back to the driver at hq_shutdown or above 256 MiB cached).  cudaMalloc/cudaFree
// synchronise the device and were measured at 2-260 ms per Circuit teardown next to a 16 GiB allocation (tools/e2e_probe.py).
cudaError_t dev_alloc(void** p, size_t bytes);
void dev_free(void* p);

// ---- bit helpers ---------------------------------------------------------------------------------
inline int popcount64(uint64_t x) { return __builtin_popcountll(x); }
inline uint64_t pdep64(uint64_t v, uint64_t mask) {
    uint64_t out = 0;
    for (uint64_t bit = 1; mask; bit <<= 1) {
        uint64_t low = mask & (~mask + 1);
        if (v & bit) out |= low;
        mask ^= low;
    }
    return out;
}

}  // namespace hq
