// Runtime + state-vector services of the device layer: process/GPU binding, allocation, |0..0> init,
// host<->device amplitude transfer, single-amplitude fetch, on-device threshold scan and norm.
// Stands in for MyGlobalVars::init (src/utils.cpp:17-60) and kernelInit / kernelDeviceToHost /
// kernelGetAmp / kernelDestroy (src/kernelSimple.cu:9-37,518-530) of the reference.
#include <map>
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "hq_internal.h"

namespace hq {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s in file %s, line %d: %04d %s", what, file, line, (int)e, cudaGetErrorString(e));
    g_err = buf;
    return HQ_ERR_CUDA;
}
Runtime& rt() {
    static Runtime r;
    return r;
}

// ---- kernels -----------------------------------------------------------------------------------------
__global__ void zero_state_kernel(double2* s, uint64_t n, int set_amp0) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        s[i] = make_double2((i == 0 && set_amp0) ? 1.0 : 0.0, 0.0);
}

// Every amplitude with |a|^2 > thresh: (index, re, im) appended to out (unordered; host sorts).
__global__ void scan_kernel(const double2* s, uint64_t n, double thresh, unsigned long long* counter, int64_t* idx_out,
                            double2* amp_out, unsigned long long cap) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double2 v = s[i];
        if (v.x * v.x + v.y * v.y > thresh) {
            const unsigned long long slot = atomicAdd(counter, 1ull);
            if (slot < cap) { idx_out[slot] = (int64_t)i; amp_out[slot] = v; }
        }
    }
}

__global__ void norm2_kernel(const double2* s, uint64_t n, double* out) {
    __shared__ double red[32];
    double acc = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double2 v = s[i];
        acc = fma(v.x, v.x, fma(v.y, v.y, acc));
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (threadIdx.x == 0) atomicAdd(out, acc);
    }
}

// P(target bit = 0) = sum |a|^2 over the amplitudes whose physical index has bit t clear (kernelMeasure / measure<>,
// src/kernelSimple.cu:482-516 of the reference).  Only that half of the state is read: index i of 2^(L-1) maps to
// lo = ((i >> t) << (t + 1)) | (i & (2^t - 1)); four independent loads in flight per thread, per-block partial sums written to
// `partial` (a fixed-order second pass adds them, so the result does not depend on atomics ordering).
__global__ void measure_kernel(const double2* __restrict__ s, uint64_t half, int t, double* __restrict__ partial) {
    __shared__ double red[32];
    const uint64_t lowmask = (1ull << t) - 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < half; i += 4 * stride) {
        double2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t j = i + (uint64_t)u * stride;
            v[u] = s[((j >> t) << (t + 1)) | (j & lowmask)];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = fma(v[u].x, v[u].x, fma(v[u].y, v[u].y, acc[u]));
    }
    for (; i < half; i += stride) {
        const double2 v = s[((i >> t) << (t + 1)) | (i & lowmask)];
        acc[0] = fma(v.x, v.x, fma(v.y, v.y, acc[0]));
    }
    double a = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x < 32) {
        a = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (threadIdx.x == 0) partial[blockIdx.x] = a;
    }
}
__global__ void sum_partials_kernel(const double* __restrict__ partial, int n, double* out) {
    __shared__ double red[32];
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a += partial[i];
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x < 32) {
        a = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (threadIdx.x == 0) *out = a;
    }
}

}  // namespace hq

using namespace hq;

extern "C" const char* hq_last_error(void) { return g_err.c_str(); }
extern "C" const char* hq_version(void) { return "hyquas_b200 0.1 (sm_100a)"; }

extern "C" int hq_device_count(int* n) {
    HQ_REQUIRE(n != nullptr, "null out pointer");
    HQ_CUDA(cudaGetDeviceCount(n));
    return HQ_OK;
}

extern "C" int hq_init(int device) {
    Runtime& r = rt();
    if (r.ready && r.device == device) return HQ_OK;
    if (r.ready) hq_shutdown();
    HQ_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    HQ_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error(std::string("hyquas_b200 needs an sm_100a GPU, found ") + prop.name);
        return HQ_ERR_UNSUPPORTED;
    }
    r.device = device;
    r.sm_count = prop.multiProcessorCount;
    HQ_CUDA(cudaStreamCreateWithFlags(&r.compute, cudaStreamNonBlocking));
    HQ_CUDA(cudaStreamCreateWithFlags(&r.comm, cudaStreamNonBlocking));
    HQ_CUDA(cudaEventCreate(&r.t0));
    HQ_CUDA(cudaEventCreate(&r.t1));
    if (const char* e = getenv("HQ_TILE_BITS")) {
        const int k = atoi(e);
        if (k >= 10 && k <= 12) r.tile_bits = k;
    }
    if (const char* e = getenv("HQ_RELAXED_REGS")) r.relaxed_regs = atoi(e) != 0;
    if (const char* e = getenv("HQ_STATE_CACHE")) r.state_cache = atoi(e) != 0;
    r.ready = true;
    return HQ_OK;
}

namespace hq {
namespace {
struct SmallPool {
    std::map<size_t, std::vector<void*>> free_lists;   // rounded size -> cached buffers
    std::map<void*, size_t> live;                      // handed-out buffer -> rounded size
    size_t cached_bytes = 0;
};
SmallPool& pool() { static SmallPool p; return p; }
size_t round_size(size_t b) { size_t r = 4096; while (r < b) r <<= 1; return r; }
}  // namespace

cudaError_t dev_alloc(void** p, size_t bytes) {
    SmallPool& sp = pool();
    const size_t r = round_size(bytes);
    auto it = sp.free_lists.find(r);
    if (it != sp.free_lists.end() && !it->second.empty()) {
        *p = it->second.back();
        it->second.pop_back();
        sp.cached_bytes -= r;
    } else {
        const cudaError_t e = cudaMalloc(p, r);
        if (e != cudaSuccess) return e;
    }
    sp.live[*p] = r;
    return cudaSuccess;
}

void dev_free(void* p) {
    if (!p) return;
    SmallPool& sp = pool();
    auto it = sp.live.find(p);
    if (it == sp.live.end()) { cudaFree(p); return; }
    const size_t r = it->second;
    sp.live.erase(it);
    if (rt().ready && sp.cached_bytes + r <= (size_t(256) << 20)) {
        // a cached buffer may be handed out again at once: work queued on it must be done (cudaFree would have waited too)
        cudaStreamSynchronize(rt().compute);
        sp.free_lists[r].push_back(p);
        sp.cached_bytes += r;
    } else {
        cudaFree(p);
    }
}

static void pool_release_all() {
    SmallPool& sp = pool();
    for (auto& kv : sp.free_lists)
        for (void* p : kv.second) cudaFree(p);
    sp.free_lists.clear();
    sp.cached_bytes = 0;
}
}  // namespace hq

// live state allocations made by hq_state_alloc -> their size (hq_state_free needs it to offer the buffer to the cache)
static std::map<void*, size_t>& rt_state_sizes() {
    static std::map<void*, size_t> m;
    return m;
}

extern "C" int hq_shutdown(void) {
    Runtime& r = rt();
    if (!r.ready) return HQ_OK;
    cudaStreamSynchronize(r.compute);
    cudaStreamSynchronize(r.comm);
    if (r.cached_state) {
        rt_state_sizes().erase(r.cached_state);
        cudaFree(r.cached_state);
    }
    hq::pool_release_all();
    cudaEventDestroy(r.t0);
    cudaEventDestroy(r.t1);
    cudaStreamDestroy(r.compute);
    cudaStreamDestroy(r.comm);
    r = Runtime();
    return HQ_OK;
}

extern "C" int hq_sync(void) {
    HQ_REQUIRE(rt().ready, "hq_init() has not been called");
    HQ_CUDA(cudaStreamSynchronize(rt().comm));
    HQ_CUDA(cudaStreamSynchronize(rt().compute));
    return HQ_OK;
}

extern "C" int hq_device_info(char* name, size_t cap, int* sm_count, size_t* total_mem) {
    HQ_REQUIRE(rt().ready, "hq_init() has not been called");
    cudaDeviceProp prop;
    HQ_CUDA(cudaGetDeviceProperties(&prop, rt().device));
    if (name && cap) snprintf(name, cap, "%s", prop.name);
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (total_mem) *total_mem = prop.totalGlobalMem;
    return HQ_OK;
}

extern "C" int hq_state_alloc(int L, void** state) {
    HQ_REQUIRE(rt().ready, "hq_init() has not been called");
    HQ_REQUIRE(state != nullptr && L >= 1 && L <= 40, "bad arguments to hq_state_alloc");
    const size_t bytes = sizeof(double2) << L;
    Runtime& r = rt();
    if (r.cached_state) {
        if (r.cached_bytes == bytes) {
            *state = r.cached_state;
            r.cached_state = nullptr;
            r.cached_bytes = 0;
            return HQ_OK;
        }
        HQ_CUDA(cudaFree(r.cached_state));   // wrong size: give it back before asking for more
        r.cached_state = nullptr;
        r.cached_bytes = 0;
    }
    HQ_CUDA(cudaMalloc(state, bytes));
    rt_state_sizes()[*state] = bytes;
    return HQ_OK;
}

extern "C" int hq_state_free(void* state) {
    if (!state) return HQ_OK;
    Runtime& r = rt();
    auto& sizes = rt_state_sizes();
    auto it = sizes.find(state);
    if (r.ready && r.state_cache && it != sizes.end()) {
        HQ_CUDA(cudaStreamSynchronize(r.compute));   // cudaFree would have synchronised; keep that guarantee
        if (r.cached_state) {
            sizes.erase(r.cached_state);
            HQ_CUDA(cudaFree(r.cached_state));
        }
        r.cached_state = state;
        r.cached_bytes = it->second;
        return HQ_OK;
    }
    if (it != sizes.end()) sizes.erase(it);
    HQ_CUDA(cudaFree(state));
    return HQ_OK;
}

extern "C" int hq_state_init(void* state, int L, int set_amp0) {
    HQ_REQUIRE(rt().ready && state != nullptr, "bad arguments to hq_state_init");
    const uint64_t n = 1ull << L;
    const int block = 256;
    const int grid = (int)std::min<uint64_t>((n + block - 1) / block, (uint64_t)rt().sm_count * 16);
    zero_state_kernel<<<grid, block, 0, rt().compute>>>(static_cast<double2*>(state), n, set_amp0);
    HQ_CUDA(cudaGetLastError());
    return HQ_OK;
}

extern "C" int hq_state_download(const void* state, int L, int64_t first, int64_t count, double* host) {
    HQ_REQUIRE(rt().ready && state && host, "bad arguments to hq_state_download");
    HQ_REQUIRE(first >= 0 && count >= 0 && (uint64_t)(first + count) <= (1ull << L), "download range outside the state");
    HQ_CUDA(cudaStreamSynchronize(rt().compute));
    HQ_CUDA(cudaMemcpy(host, static_cast<const double2*>(state) + first, (size_t)count * sizeof(double2), cudaMemcpyDeviceToHost));
    return HQ_OK;
}

extern "C" int hq_state_upload(void* state, int L, int64_t first, int64_t count, const double* host) {
    HQ_REQUIRE(rt().ready && state && host, "bad arguments to hq_state_upload");
    HQ_REQUIRE(first >= 0 && count >= 0 && (uint64_t)(first + count) <= (1ull << L), "upload range outside the state");
    HQ_CUDA(cudaStreamSynchronize(rt().compute));
    HQ_CUDA(cudaMemcpy(static_cast<double2*>(state) + first, host, (size_t)count * sizeof(double2), cudaMemcpyHostToDevice));
    return HQ_OK;
}

extern "C" int hq_amp_fetch(const void* state, int64_t idx, double out[2]) {
    HQ_REQUIRE(rt().ready && state && out && idx >= 0, "bad arguments to hq_amp_fetch");
    HQ_CUDA(cudaStreamSynchronize(rt().compute));
    HQ_CUDA(cudaMemcpy(out, static_cast<const double2*>(state) + idx, sizeof(double2), cudaMemcpyDeviceToHost));
    return HQ_OK;
}

extern "C" int hq_dump_scan(const void* state, int L, double thresh, int64_t* idx_out, double* amp_out, int64_t cap, int64_t* found) {
    HQ_REQUIRE(rt().ready && state && idx_out && amp_out && found && cap > 0, "bad arguments to hq_dump_scan");
    const uint64_t n = 1ull << L;
    unsigned long long* d_cnt = nullptr;
    int64_t* d_idx = nullptr;
    double2* d_amp = nullptr;
    HQ_CUDA(dev_alloc(reinterpret_cast<void**>(&d_cnt), 8));
    HQ_CUDA(dev_alloc(reinterpret_cast<void**>(&d_idx), (size_t)cap * 8));
    HQ_CUDA(dev_alloc(reinterpret_cast<void**>(&d_amp), (size_t)cap * 16));
    HQ_CUDA(cudaMemsetAsync(d_cnt, 0, 8, rt().compute));
    const int block = 256;
    const int grid = (int)std::min<uint64_t>((n + block - 1) / block, (uint64_t)rt().sm_count * 16);
    scan_kernel<<<grid, block, 0, rt().compute>>>(static_cast<const double2*>(state), n, thresh, d_cnt, d_idx, d_amp,
                                                    (unsigned long long)cap);
    HQ_CUDA(cudaGetLastError());
    unsigned long long cnt = 0;
    HQ_CUDA(cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, rt().compute));
    HQ_CUDA(cudaStreamSynchronize(rt().compute));
    const int64_t m = (int64_t)std::min<unsigned long long>(cnt, (unsigned long long)cap);
    std::vector<int64_t> hidx(m);
    std::vector<double2> hamp(m);
    if (m) {
        HQ_CUDA(cudaMemcpy(hidx.data(), d_idx, (size_t)m * 8, cudaMemcpyDeviceToHost));
        HQ_CUDA(cudaMemcpy(hamp.data(), d_amp, (size_t)m * 16, cudaMemcpyDeviceToHost));
    }
    dev_free(d_cnt); dev_free(d_idx); dev_free(d_amp);
    std::vector<int64_t> order(m);
    for (int64_t i = 0; i < m; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return hidx[a] < hidx[b]; });
    for (int64_t i = 0; i < m; ++i) {
        idx_out[i] = hidx[order[i]];
        amp_out[2 * i] = hamp[order[i]].x;
        amp_out[2 * i + 1] = hamp[order[i]].y;
    }
    *found = (int64_t)cnt;
    return HQ_OK;
}

extern "C" int hq_state_norm2(const void* state, int L, double* out) {
    HQ_REQUIRE(rt().ready && state && out, "bad arguments to hq_state_norm2");
    const uint64_t n = 1ull << L;
    double* d = nullptr;
    HQ_CUDA(dev_alloc(reinterpret_cast<void**>(&d), 8));
    HQ_CUDA(cudaMemsetAsync(d, 0, 8, rt().compute));
    const int block = 256;
    const int grid = (int)std::min<uint64_t>((n + block - 1) / block, (uint64_t)rt().sm_count * 8);
    norm2_kernel<<<grid, block, 0, rt().compute>>>(static_cast<const double2*>(state), n, d);
    HQ_CUDA(cudaGetLastError());
    HQ_CUDA(cudaMemcpyAsync(out, d, 8, cudaMemcpyDeviceToHost, rt().compute));
    HQ_CUDA(cudaStreamSynchronize(rt().compute));
    dev_free(d);
    return HQ_OK;
}

// Probability that physical local bit `target_bit` reads 0 (reference: kernelMeasure, src/kernel.h:14).  HBM-bound: reads
// 16 * 2^(L-1) bytes (the "bit clear" half; whole 32-byte sectors for target_bit 0, i.e. 16 * 2^L there).
extern "C" int hq_state_measure(const void* state, int L, int target_bit, double* p0) {
    HQ_REQUIRE(rt().ready && state && p0 && L >= 1 && target_bit >= 0 && target_bit < L, "bad arguments to hq_state_measure");
    const uint64_t half = 1ull << (L - 1);
    const int block = 256;
    const int grid = (int)std::min<uint64_t>((half + block - 1) / block, (uint64_t)rt().sm_count * 8);
    double* d = nullptr;
    HQ_CUDA(dev_alloc(reinterpret_cast<void**>(&d), (size_t)(grid + 1) * 8));
    measure_kernel<<<grid, block, 0, rt().compute>>>(static_cast<const double2*>(state), half, target_bit, d + 1);
    HQ_CUDA(cudaGetLastError());
    sum_partials_kernel<<<1, 256, 0, rt().compute>>>(d + 1, grid, d);
    HQ_CUDA(cudaGetLastError());
    HQ_CUDA(cudaMemcpyAsync(p0, d, 8, cudaMemcpyDeviceToHost, rt().compute));
    HQ_CUDA(cudaStreamSynchronize(rt().compute));
    dev_free(d);
    return HQ_OK;
}

extern "C" int hq_timer_start(void) {
    HQ_REQUIRE(rt().ready, "hq_init() has not been called");
    HQ_CUDA(cudaEventRecord(rt().t0, rt().compute));
    return HQ_OK;
}

extern "C" int hq_timer_stop_ms(float* ms) {
    HQ_REQUIRE(rt().ready && ms, "bad arguments to hq_timer_stop_ms");
    HQ_CUDA(cudaEventRecord(rt().t1, rt().compute));
    HQ_CUDA(cudaEventSynchronize(rt().t1));
    HQ_CUDA(cudaEventElapsedTime(ms, rt().t0, rt().t1));
    return HQ_OK;
}
