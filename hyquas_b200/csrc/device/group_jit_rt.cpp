// NVRTC runtime of the per-group specialised kernels: compile for sm_100a, cache (memory + disk), load, launch.
//
// No link-time dependency on NVRTC or libcuda: NVRTC is dlopen()ed and the few driver entry points come from
// cudaGetDriverEntryPoint, so the library still loads on a machine without either (CPU tests; the product path then keeps
// the interpreter kernel of group_kernel.cu -- still this library's CUDA path, never a CPU fallback).
//
// Cache: key = 128-bit hash of (the plan's identity bytes, compile options, NVRTC version, emitter version, kernel skeleton
// text) -- the source itself is only emitted on a miss.  Hits in the process-wide map cost nothing; hits on disk ($HQ_JIT_CACHE,
// else $XDG_CACHE_HOME/hyquas_b200/jit, else ~/.cache/hyquas_b200/jit, else /tmp/hyquas_b200_jit_<uid>) cost a file read +
// cuModuleLoadData; misses cost one NVRTC compile (0.1-3 s for 20-150 gates), done on all host cores when a whole schedule is
// prepared at once (jit_precompile).  A cached file the driver refuses to load is deleted and compiled again.  The role of the reference's process-global cuTT plan cache
// (src/schedule.cpp:723-783), for kernels instead of transpose plans.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <omp.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "group_jit.h"
#include "hq_internal.h"

namespace hq {

struct JitKernel {
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
    size_t smem_set = 0;
    int regs = 0;
    int spill_bytes = 0;
    int refs = 0;             // plans holding this kernel (jit_get / jit_release)
    uint64_t last_use = 0;    // for eviction: loaded modules are capped, the least recently fetched unreferenced one goes first
};

namespace {

struct Nvrtc {
    void* h = nullptr;
    bool tried = false;
    decltype(&nvrtcCreateProgram) create = nullptr;
    decltype(&nvrtcCompileProgram) compile = nullptr;
    decltype(&nvrtcGetCUBINSize) cubin_size = nullptr;
    decltype(&nvrtcGetCUBIN) cubin = nullptr;
    decltype(&nvrtcGetProgramLogSize) log_size = nullptr;
    decltype(&nvrtcGetProgramLog) log = nullptr;
    decltype(&nvrtcDestroyProgram) destroy = nullptr;
    decltype(&nvrtcVersion) version = nullptr;
    int major = 0, minor = 0;
    bool ok() const { return h != nullptr; }
};

Nvrtc& nvrtc() {
    static Nvrtc n;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (n.tried) return n;
    n.tried = true;
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so.13"};
    for (const char* nm : names) {
        n.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (n.h) break;
    }
    if (!n.h) return n;
#define HQ_SYM(field, name) n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.h, name)); if (!n.field) { dlclose(n.h); n.h = nullptr; return n; }
    HQ_SYM(create, "nvrtcCreateProgram")
    HQ_SYM(compile, "nvrtcCompileProgram")
    HQ_SYM(cubin_size, "nvrtcGetCUBINSize")
    HQ_SYM(cubin, "nvrtcGetCUBIN")
    HQ_SYM(log_size, "nvrtcGetProgramLogSize")
    HQ_SYM(log, "nvrtcGetProgramLog")
    HQ_SYM(destroy, "nvrtcDestroyProgram")
    HQ_SYM(version, "nvrtcVersion")
#undef HQ_SYM
    n.version(&n.major, &n.minor);
    return n;
}

struct Driver {
    bool tried = false, good = false;
    CUresult (*moduleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*moduleUnload)(CUmodule) = nullptr;
    CUresult (*moduleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*funcSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*funcGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
    CUresult (*launchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*occupancy)(int*, CUfunction, int, size_t) = nullptr;
    CUresult (*getErrorString)(CUresult, const char**) = nullptr;
};

Driver& driver() {
    static Driver d;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (d.tried) return d;
    d.tried = true;
    bool all = true;
    auto get = [&](const char* name, void** fp) {
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint(name, fp, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !*fp) all = false;
    };
    get("cuModuleLoadData", reinterpret_cast<void**>(&d.moduleLoadData));
    get("cuModuleUnload", reinterpret_cast<void**>(&d.moduleUnload));
    get("cuModuleGetFunction", reinterpret_cast<void**>(&d.moduleGetFunction));
    get("cuFuncSetAttribute", reinterpret_cast<void**>(&d.funcSetAttribute));
    get("cuFuncGetAttribute", reinterpret_cast<void**>(&d.funcGetAttribute));
    get("cuLaunchKernel", reinterpret_cast<void**>(&d.launchKernel));
    get("cuOccupancyMaxActiveBlocksPerMultiprocessor", reinterpret_cast<void**>(&d.occupancy));
    get("cuGetErrorString", reinterpret_cast<void**>(&d.getErrorString));
    (void)cudaGetLastError();
    d.good = all;
    return d;
}

struct Key {
    uint64_t a, b;
    bool operator<(const Key& o) const { return a != o.a ? a < o.a : b < o.b; }
};
uint64_t fnv(const std::string& s, uint64_t h) {
    for (unsigned char c : s) { h ^= c; h *= 0x100000001b3ull; }
    return h;
}
const char* kEmitterVersion = "hqjit4";   // bump when group_jit.cpp changes what it emits for the same plan
const char* kOptions[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-default-device", "--ptxas-options=-v", "-lineinfo"};
// `what` identifies the kernel: the bytes of the plan it is generated from (see plan_identity() in group_kernel.cu), so that a
// cache hit costs a hash of ~20 KB and no source emission at all.  Compiler version, options and the emitter's version salt
// are part of the key.
Key key_of(const std::string& what) {
    static const std::string salt = [] {
        std::string t = std::string(kEmitterVersion) + "|" + std::to_string(nvrtc().major) + "." + std::to_string(nvrtc().minor);
        for (const char* o : kOptions) t += std::string("|") + o;
        // the skeleton every kernel is generated around: an edit there invalidates old cubins without a version bump
        t += "|" + std::to_string(fnv(jit_device_prologue(), 0xcbf29ce484222325ull));
        return t;
    }();
    Key k;
    k.a = fnv(what, fnv(salt, 0xcbf29ce484222325ull));
    k.b = fnv(what, fnv(salt, 0x9e3779b97f4a7c15ull) ^ what.size());
    return k;
}

struct Cache {
    std::mutex mu;
    std::map<Key, JitKernel*> loaded;
    std::map<Key, std::vector<char>> cubins;   // compiled or read from disk, not yet loaded (jit_precompile)
    int compiled = 0, disk_hits = 0;
    uint64_t clock = 0;
    size_t max_loaded = 512;   // loaded modules kept (HQ_JIT_MAX_LOADED)
    double compile_seconds = 0;
    std::string dir;           // set once by cache_dir()
};
Cache& cache() { static Cache c; return c; }

void init_cache_dir(Cache& c);
const std::string& cache_dir() {
    static std::once_flag once;
    Cache& c = cache();
    std::call_once(once, [&] { init_cache_dir(c); });
    return c.dir;
}
void init_cache_dir(Cache& c) {
    if (const char* e = getenv("HQ_JIT_MAX_LOADED")) c.max_loaded = (size_t)std::max(8, atoi(e));
    std::string base;
    if (const char* e = getenv("HQ_JIT_CACHE")) base = e;
    else if (const char* x = getenv("XDG_CACHE_HOME")) base = std::string(x) + "/hyquas_b200/jit";
    else if (const char* h = getenv("HOME")) base = std::string(h) + "/.cache/hyquas_b200/jit";
    else base = "/tmp/hyquas_b200_jit_" + std::to_string((int)getuid());
    if (base == "off" || base == "0") return;   // memory cache only
    std::string path;
    for (size_t i = 1; i <= base.size(); ++i)
        if (i == base.size() || base[i] == '/') { path = base.substr(0, i); mkdir(path.c_str(), 0700); }
    struct stat st;
    if (stat(base.c_str(), &st) == 0 && S_ISDIR(st.st_mode) && access(base.c_str(), W_OK) == 0) c.dir = base;
}

std::string file_of(const Key& k) {
    char buf[64];
    snprintf(buf, sizeof(buf), "/%016llx%016llx.cubin", (unsigned long long)k.a, (unsigned long long)k.b);
    return cache_dir() + buf;
}

bool read_disk(const Key& k, std::vector<char>& out) {
    if (cache_dir().empty()) return false;
    std::ifstream in(file_of(k), std::ios::binary | std::ios::ate);
    if (!in) return false;
    const std::streamsize n = in.tellg();
    if (n < 64) return false;
    out.resize((size_t)n);
    in.seekg(0);
    in.read(out.data(), n);
    return (bool)in && std::memcmp(out.data(), "\x7f" "ELF", 4) == 0;
}

void write_disk(const Key& k, const std::vector<char>& cubin) {
    if (cache_dir().empty()) return;
    const std::string dst = file_of(k), tmp = dst + ".tmp" + std::to_string((long)getpid()) + "_" + std::to_string(omp_get_thread_num());
    {
        std::ofstream out(tmp, std::ios::binary);
        if (!out) return;
        out.write(cubin.data(), (std::streamsize)cubin.size());
        if (!out) { unlink(tmp.c_str()); return; }
    }
    if (rename(tmp.c_str(), dst.c_str()) != 0) unlink(tmp.c_str());
}

bool verbose() { static const bool v = getenv("HQ_JIT_VERBOSE") != nullptr; return v; }

// NVRTC is thread-safe: called concurrently from jit_precompile
bool compile(const std::string& src, std::vector<char>& cubin, std::string* why, std::string* log_out) {
    Nvrtc& n = nvrtc();
    if (!n.ok()) { if (why) *why = "libnvrtc not found"; return false; }
    nvrtcProgram prog = nullptr;
    if (n.create(&prog, src.c_str(), "hq_group_jit.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) { if (why) *why = "nvrtcCreateProgram failed"; return false; }
    const nvrtcResult rc = n.compile(prog, (int)(sizeof(kOptions) / sizeof(kOptions[0])), kOptions);
    size_t ls = 0;
    n.log_size(prog, &ls);
    std::string log(ls, '\0');
    if (ls > 1) n.log(prog, &log[0]);
    if (log_out) *log_out = log;
    if (rc != NVRTC_SUCCESS) {
        if (why) *why = "NVRTC compile failed: " + log.substr(0, 2000);
        n.destroy(&prog);
        return false;
    }
    size_t cs = 0;
    n.cubin_size(prog, &cs);
    cubin.resize(cs);
    n.cubin(prog, cubin.data());
    n.destroy(&prog);
    return cs > 0;
}

bool obtain_cubin(const Key& k, const std::string& src, std::vector<char>& cubin, std::string* why) {
    Cache& c = cache();
    {
        std::lock_guard<std::mutex> lock(c.mu);
        auto it = c.cubins.find(k);
        if (it != c.cubins.end()) { cubin = it->second; return true; }
    }
    if (read_disk(k, cubin)) {
        std::lock_guard<std::mutex> lock(c.mu);
        ++c.disk_hits;
        return true;
    }
    const auto t0 = std::chrono::steady_clock::now();
    std::string log;
    if (!compile(src, cubin, why, &log)) return false;
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (verbose()) fprintf(stderr, "[hq jit] compiled %zu-byte source in %.2f s -> %zu-byte cubin\n%s\n", src.size(), dt, cubin.size(), log.c_str());
    write_disk(k, cubin);
    std::lock_guard<std::mutex> lock(c.mu);
    ++c.compiled;
    c.compile_seconds += dt;
    return true;
}

}  // namespace

bool jit_enabled() {
    static const bool on = [] {
        const char* e = getenv("HQ_JIT");
        return !(e && (e[0] == '0' || e[0] == 'n' || e[0] == 'N'));
    }();
    return on;
}

bool jit_cached(const std::string& identity) {
    if (!nvrtc().ok()) return false;
    Cache& c = cache();
    const Key k = key_of(identity);
    std::lock_guard<std::mutex> lock(c.mu);
    return c.loaded.count(k) || c.cubins.count(k);
}

// identities[i] names kernel i; emit(i) produces its source when (and only when) neither the memory nor the disk cache has it.
void jit_precompile(const std::string* identities, int n, const std::function<std::string(int)>& emit) {
    if (!nvrtc().ok()) return;
    cache_dir();
    Cache& c = cache();
    std::vector<int> todo;
    std::vector<Key> keys(n);
    for (int i = 0; i < n; ++i) {
        if (identities[i].empty()) continue;
        keys[i] = key_of(identities[i]);
        std::lock_guard<std::mutex> lock(c.mu);
        bool dup = c.loaded.count(keys[i]) || c.cubins.count(keys[i]);
        for (int j : todo) if (!(keys[j] < keys[i]) && !(keys[i] < keys[j])) dup = true;
        if (!dup) todo.push_back(i);
    }
    const int m = (int)todo.size();
    #pragma omp parallel for schedule(dynamic, 1) if (m > 1)
    for (int j = 0; j < m; ++j) {
        const int i = todo[j];
        std::vector<char> cubin;
        if (!read_disk(keys[i], cubin)) {
            const std::string src = emit(i);
            if (src.empty() || !obtain_cubin(keys[i], src, cubin, nullptr)) continue;
        } else {
            std::lock_guard<std::mutex> lock(c.mu);
            ++c.disk_hits;
        }
        std::lock_guard<std::mutex> lock(c.mu);
        c.cubins[keys[i]] = std::move(cubin);
    }
}

JitKernel* jit_get(const std::string& identity, size_t dynamic_smem, const std::function<std::string()>& emit, std::string* why) {
    Driver& d = driver();
    if (!d.good) { if (why) *why = "CUDA driver entry points unavailable"; return nullptr; }
    if (!nvrtc().ok()) { if (why) *why = "libnvrtc not found"; return nullptr; }
    cache_dir();
    Cache& c = cache();
    const Key k = key_of(identity);
    {
        std::lock_guard<std::mutex> lock(c.mu);
        auto it = c.loaded.find(k);
        if (it != c.loaded.end()) { ++it->second->refs; it->second->last_use = ++c.clock; return it->second; }
    }
    std::vector<char> cubin;
    bool have = false, from_disk = false;
    {
        std::lock_guard<std::mutex> lock(c.mu);
        auto it = c.cubins.find(k);
        if (it != c.cubins.end()) { cubin = it->second; have = true; }
    }
    if (!have && read_disk(k, cubin)) {
        std::lock_guard<std::mutex> lock(c.mu);
        ++c.disk_hits;
        have = from_disk = true;
    }
    auto build = [&]() {
        const std::string src = emit();
        if (src.empty()) { if (why) *why = "the emitter rejected this plan"; return false; }
        return obtain_cubin(k, src, cubin, why);
    };
    if (!have && !build()) return nullptr;
    auto fail = [&](CUresult r, const char* what) {
        const char* s = nullptr;
        d.getErrorString(r, &s);
        if (why) *why = std::string(what) + ": " + (s ? s : "?");
        return (JitKernel*)nullptr;
    };
    cudaSetDevice(rt().device);   // make the primary context current on this thread
    auto* jk = new JitKernel();
    CUresult r = d.moduleLoadData(&jk->mod, cubin.data());
    if (r != CUDA_SUCCESS && from_disk) {   // a damaged or foreign file in the cache directory: drop it and compile
        unlink(file_of(k).c_str());
        cubin.clear();
        if (!build()) { delete jk; return nullptr; }
        r = d.moduleLoadData(&jk->mod, cubin.data());
    }
    if (r != CUDA_SUCCESS) { delete jk; return fail(r, "cuModuleLoadData"); }
    r = d.moduleGetFunction(&jk->fn, jk->mod, "hq_group_jit");
    if (r != CUDA_SUCCESS) { d.moduleUnload(jk->mod); delete jk; return fail(r, "cuModuleGetFunction"); }
    r = d.funcSetAttribute(jk->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)dynamic_smem);
    if (r != CUDA_SUCCESS) { d.moduleUnload(jk->mod); delete jk; return fail(r, "cuFuncSetAttribute"); }
    jk->smem_set = dynamic_smem;
    d.funcGetAttribute(&jk->regs, CU_FUNC_ATTRIBUTE_NUM_REGS, jk->fn);
    d.funcGetAttribute(&jk->spill_bytes, CU_FUNC_ATTRIBUTE_LOCAL_SIZE_BYTES, jk->fn);
    std::lock_guard<std::mutex> lock(c.mu);
    c.cubins.erase(k);
    auto raced = c.loaded.find(k);
    if (raced != c.loaded.end()) {   // another host thread loaded the same kernel meanwhile: keep theirs
        d.moduleUnload(jk->mod);
        delete jk;
        ++raced->second->refs;
        raced->second->last_use = ++c.clock;
        return raced->second;
    }
    jk->refs = 1;
    jk->last_use = ++c.clock;
    c.loaded[k] = jk;
    // a long-lived process that keeps compiling new circuits must not accumulate modules without bound
    while (c.loaded.size() > c.max_loaded) {
        auto victim = c.loaded.end();
        for (auto it = c.loaded.begin(); it != c.loaded.end(); ++it)
            if (it->second->refs == 0 && (victim == c.loaded.end() || it->second->last_use < victim->second->last_use)) victim = it;
        if (victim == c.loaded.end()) break;   // everything is in use
        d.moduleUnload(victim->second->mod);
        delete victim->second;
        c.loaded.erase(victim);
    }
    return jk;
}

void jit_release(JitKernel* k) {
    if (!k) return;
    Cache& c = cache();
    std::lock_guard<std::mutex> lock(c.mu);
    if (k->refs > 0) --k->refs;
}

int jit_launch(JitKernel* k, int grid, int block, size_t smem, void* stream, void* state, int amp0) {
    Driver& d = driver();
    void* args[] = {&state, &amp0};
    const CUresult r = d.launchKernel(k->fn, (unsigned)grid, 1, 1, (unsigned)block, 1, 1, (unsigned)smem, static_cast<CUstream>(stream), args, nullptr);
    if (r != CUDA_SUCCESS) {
        const char* s = nullptr;
        d.getErrorString(r, &s);
        set_error(std::string("cuLaunchKernel(hq_group_jit): ") + (s ? s : "?"));
        return HQ_ERR_CUDA;
    }
    return HQ_OK;
}

int jit_max_blocks_per_sm(JitKernel* k, int block, size_t smem) {
    int n = 0;
    if (driver().occupancy(&n, k->fn, block, smem) != CUDA_SUCCESS) return 1;
    return n < 1 ? 1 : n;
}

void jit_kernel_info(JitKernel* k, int* regs, int* spill) { if (regs) *regs = k->regs; if (spill) *spill = k->spill_bytes; }

void jit_stats(int* kernels, int* compiled, int* disk_hits, double* compile_seconds) {
    Cache& c = cache();
    std::lock_guard<std::mutex> lock(c.mu);
    if (kernels) *kernels = (int)c.loaded.size();
    if (compiled) *compiled = c.compiled;
    if (disk_hits) *disk_hits = c.disk_hits;
    if (compile_seconds) *compile_seconds = c.compile_seconds;
}

}  // namespace hq

// Offline use of the compiler (tests, tools): source -> cubin bytes written to `path`.  No GPU needed.
extern "C" int hq_debug_jit_compile_to_file(const char* source, const char* path, char* log, size_t log_cap) {
    std::vector<char> cubin;
    std::string why, lg;
    const bool ok = hq::compile(source, cubin, &why, &lg);
    if (log && log_cap) { snprintf(log, log_cap, "%s", ok ? lg.c_str() : why.c_str()); }
    if (!ok) { hq::set_error(why); return HQ_ERR_UNSUPPORTED; }
    std::ofstream out(path, std::ios::binary);
    out.write(cubin.data(), (std::streamsize)cubin.size());
    return out ? HQ_OK : HQ_ERR_ARG;
}

// The cache layer without a GPU (tests): the kernel named `identity` goes through jit_precompile with `source` as its text;
// *compiled / *disk_hits are the process totals afterwards.  What jit_get adds on top is the module load.
extern "C" int hq_debug_jit_cache_probe(const char* identity, const char* source, int* compiled, int* disk_hits) {
    if (!identity || !source) { hq::set_error("null argument"); return HQ_ERR_ARG; }
    if (!hq::nvrtc().ok()) { hq::set_error("libnvrtc not found"); return HQ_ERR_UNSUPPORTED; }
    const std::string id = identity, src = source;
    hq::jit_precompile(&id, 1, [&](int) { return src; });
    hq::jit_stats(nullptr, compiled, disk_hits, nullptr);
    return hq::jit_cached(id) ? HQ_OK : HQ_ERR_UNSUPPORTED;
}

// where this library keeps files between runs (compiled kernels, the partitioner's search results); empty = disk caching is off
extern "C" int hq_cache_dir(char* out, size_t cap) {
    if (!out || cap == 0) { hq::set_error("null argument"); return HQ_ERR_ARG; }
    const std::string& d = hq::cache_dir();
    if (d.size() + 1 > cap) { hq::set_error("buffer too small"); return HQ_ERR_ARG; }
    std::memcpy(out, d.c_str(), d.size() + 1);
    return HQ_OK;
}

// 1 when gate groups will run as specialised kernels: HQ_JIT not 0 and NVRTC loadable (no GPU needed to answer)
extern "C" int hq_jit_available(int* yes) {
    if (yes) *yes = hq::jit_enabled() && hq::nvrtc().ok();
    return HQ_OK;
}

extern "C" int hq_jit_stats(int* kernels_loaded, int* compiled, int* disk_hits, double* compile_seconds) {
    hq::jit_stats(kernels_loaded, compiled, disk_hits, compile_seconds);
    return HQ_OK;
}
