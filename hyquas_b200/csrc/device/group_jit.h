// Per-group specialised gate-group kernels: source emitter (group_jit.cpp) and NVRTC runtime + cache (group_jit_rt.cpp).
#pragma once
#include <cstddef>
#include <functional>
#include <string>

struct hq_group_plan;

namespace hq {

// CUDA C++ source of the kernel `hq_group_jit(double2* state)` that applies exactly this plan (host = false), or of the
// serial C++ function `hq_group_jit_host(double* state)` with the same arithmetic text (host = true; CPU tests only).
// Empty string: the emitter does not handle this plan (the interpreter kernel does).
std::string jit_emit_source(const hq_group_plan& plan, bool host, bool zero_input = false);
double jit_fp64_per_amp(const hq_group_plan& plan);
int jit_min_blocks(int K);
int jit_l2_prefetch_slots();
size_t jit_smem_bytes(int K);   // dynamic shared memory of the kernel: three tile buffers + mbarriers
const char* jit_device_prologue();
const char* jit_device_epilogue();
const char* jit_host_prologue();
const char* jit_host_epilogue();

struct JitKernel;   // one loaded cubin
// Compile (or fetch from the in-memory / on-disk cache) and load.  Returns nullptr and sets `why` when NVRTC or the driver
// entry points are unavailable or the compile fails; the caller then keeps the interpreter kernel.
// `identity` = the bytes that determine the kernel (the plan); `emit` is only called on a cache miss.
JitKernel* jit_get(const std::string& identity, size_t dynamic_smem, const std::function<std::string()>& emit, std::string* why);
// Compile a batch on all host cores (cold cache), without loading; a later jit_get() finds them cached.
void jit_precompile(const std::string* identities, int n, const std::function<std::string(int)>& emit);
bool jit_cached(const std::string& identity);
void jit_release(JitKernel* k);   // the plan that fetched it is gone (unreferenced kernels may be evicted, 512 stay loaded)
int jit_launch(JitKernel* k, int grid, int block, size_t smem, void* stream, void* state, int amp0 = 0);
int jit_max_blocks_per_sm(JitKernel* k, int block, size_t smem);
void jit_stats(int* kernels, int* compiled, int* disk_hits, double* compile_seconds);
bool jit_enabled();

}  // namespace hq
