// Gate-group kernel for sm_100a (the OShareMem-class path).
//
// What it computes is what the reference's run<128> computes (src/kernelOpt.cu:388-433): for every
// tile of the local state (the amplitudes whose tile_mask bits vary while all other bits are fixed)
// apply, in order, a list of single-/controlled-qubit gates, in place.  How it does it is new:
//
//   * persistent CTAs, one producer warp + NT consumer threads; the producer streams tiles into a
//     ring of shared-memory buffers with 1-D TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx),
//     one copy per contiguous run of the tile, so HBM reads are asynchronous and never touch registers;
//   * the host splits the gate list into ROUNDS.  In a round every consumer thread holds 16 amplitudes
//     (4 "register qubits") in registers and applies all gates of the round there: no shared-memory
//     traffic and no barrier per gate (the reference does one SMEM read-modify-write plus a
//     __syncthreads() per gate, kernelOpt.cu:214-386).  Between rounds the tile is re-laid-out through
//     shared memory (XOR-swizzled, conflict-free 128-bit accesses) to change the register qubits;
//   * controls and diagonal targets may sit on register bits, thread bits, or bits outside the tile
//     (the reference's "block bits"); they become predicate masks over the physical index;
//   * the last round writes registers straight back to HBM with 128-bit stores.
//
// Algorithmic traffic: 32 bytes per amplitude per launch (16 read + 16 written), independent of the
// number of gates.
#include <algorithm>
#include <cstring>
#include <vector>

#include "hq_internal.h"
#include "group_plan.h"

namespace hq {

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
template <int NT>
__device__ __forceinline__ void consumer_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
}

// ---- in-register gate arithmetic -----------------------------------------------------------------
#define HQ_PAIR_LOOP(TB)                                                     \
    _Pragma("unroll") for (int p = 0; p < R / 2; ++p) {                      \
        const int lo = ((p >> (TB)) << ((TB) + 1)) | (p & ((1 << (TB)) - 1)); \
        const int hi = lo | (1 << (TB));                                     \
        if ((lo & creg) == creg)

template <int TB>
__device__ __forceinline__ void op_gen(double2 (&a)[R], const DevOp& o, uint32_t creg) {
    const double r00 = o.m[0], i00 = o.m[1], r01 = o.m[2], i01 = o.m[3];
    const double r10 = o.m[4], i10 = o.m[5], r11 = o.m[6], i11 = o.m[7];
    HQ_PAIR_LOOP(TB) {
        const double2 x = a[lo], y = a[hi];
        a[lo].x = fma(-i01, y.y, fma(r01, y.x, fma(-i00, x.y, r00 * x.x)));
        a[lo].y = fma(r01, y.y, fma(i01, y.x, fma(r00, x.y, i00 * x.x)));
        a[hi].x = fma(-i11, y.y, fma(r11, y.x, fma(-i10, x.y, r10 * x.x)));
        a[hi].y = fma(r11, y.y, fma(i11, y.x, fma(r10, x.y, i10 * x.x)));
    }}
}
template <int TB>
__device__ __forceinline__ void op_real(double2 (&a)[R], const DevOp& o, uint32_t creg) {
    const double r00 = o.m[0], r01 = o.m[2], r10 = o.m[4], r11 = o.m[6];
    HQ_PAIR_LOOP(TB) {
        const double2 x = a[lo], y = a[hi];
        a[lo].x = fma(r01, y.x, r00 * x.x);
        a[lo].y = fma(r01, y.y, r00 * x.y);
        a[hi].x = fma(r11, y.x, r10 * x.x);
        a[hi].y = fma(r11, y.y, r10 * x.y);
    }}
}
template <int TB>
__device__ __forceinline__ void op_rxl(double2 (&a)[R], const DevOp& o, uint32_t creg) {
    const double r00 = o.m[0], i01 = o.m[3], i10 = o.m[5], r11 = o.m[6];
    HQ_PAIR_LOOP(TB) {
        const double2 x = a[lo], y = a[hi];
        a[lo].x = fma(-i01, y.y, r00 * x.x);
        a[lo].y = fma(i01, y.x, r00 * x.y);
        a[hi].x = fma(-i10, x.y, r11 * y.x);
        a[hi].y = fma(i10, x.x, r11 * y.y);
    }}
}
template <int TB>
__device__ __forceinline__ void op_swap(double2 (&a)[R], uint32_t creg) {
    HQ_PAIR_LOOP(TB) {
        const double2 x = a[lo];
        a[lo] = a[hi];
        a[hi] = x;
    }}
}
template <int TB>
__device__ __forceinline__ void op_yl(double2 (&a)[R], uint32_t creg) {
    HQ_PAIR_LOOP(TB) {
        const double2 x = a[lo], y = a[hi];
        a[lo] = make_double2(y.y, -y.x);
        a[hi] = make_double2(-x.y, x.x);
    }}
}
template <int TB>
__device__ __forceinline__ void op_diag_r(double2 (&a)[R], const DevOp& o, uint32_t creg) {
    const double r0 = o.m[0], i0 = o.m[1], r1 = o.m[6], i1 = o.m[7];
    const bool skip_lo = o.flags & 1u;
    HQ_PAIR_LOOP(TB) {
        if (!skip_lo) {
            const double2 x = a[lo];
            a[lo].x = fma(-i0, x.y, r0 * x.x);
            a[lo].y = fma(i0, x.x, r0 * x.y);
        }
        const double2 y = a[hi];
        a[hi].x = fma(-i1, y.y, r1 * y.x);
        a[hi].y = fma(i1, y.x, r1 * y.y);
    }}
}

#define HQ_TB_SWITCH(CALL)          \
    switch (o.tbit) {               \
        case 0: CALL(0); break;     \
        case 1: CALL(1); break;     \
        case 2: CALL(2); break;     \
        default: CALL(3); break;    \
    }

__device__ __forceinline__ void apply_op(double2 (&a)[R], const DevOp& o, uint64_t phys) {
    if ((phys & o.cphys) != o.cphys) return;
    const uint32_t creg = o.creg;
    switch (o.kind) {
        case OP_GEN: {
#define C_(TB) op_gen<TB>(a, o, creg)
            HQ_TB_SWITCH(C_)
#undef C_
            break;
        }
        case OP_REAL: {
#define C_(TB) op_real<TB>(a, o, creg)
            HQ_TB_SWITCH(C_)
#undef C_
            break;
        }
        case OP_RXL: {
#define C_(TB) op_rxl<TB>(a, o, creg)
            HQ_TB_SWITCH(C_)
#undef C_
            break;
        }
        case OP_SWAP: {
#define C_(TB) op_swap<TB>(a, creg)
            HQ_TB_SWITCH(C_)
#undef C_
            break;
        }
        case OP_YL: {
#define C_(TB) op_yl<TB>(a, creg)
            HQ_TB_SWITCH(C_)
#undef C_
            break;
        }
        case OP_DIAG_R: {
#define C_(TB) op_diag_r<TB>(a, o, creg)
            HQ_TB_SWITCH(C_)
#undef C_
            break;
        }
        default: {  // OP_DIAG_T
            const bool hi = (o.tphys == 0) || (phys & o.tphys);
            if (!hi && (o.flags & 1u)) return;
            const double fr = hi ? o.m[6] : o.m[0], fi = hi ? o.m[7] : o.m[1];
#pragma unroll
            for (int i = 0; i < R; ++i) {
                if ((i & creg) == creg) {
                    const double2 x = a[i];
                    a[i].x = fma(-fi, x.y, fr * x.x);
                    a[i].y = fma(fi, x.x, fr * x.y);
                }
            }
            break;
        }
    }
}

__device__ __forceinline__ DevOp load_op(const DevOp* p) {
    DevOp o;
    const uint4* s = reinterpret_cast<const uint4*>(p);
    uint4* d = reinterpret_cast<uint4*>(&o);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(DevOp) / 16); ++i) d[i] = __ldg(s + i);
    return o;
}

// ---- the kernel ----------------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__((1 << (K - RBITS)) + 32, 1) group_kernel(const __grid_constant__ GroupParams P) {
    constexpr int NT = 1 << (K - RBITS);
    constexpr int TILE = 1 << K;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* tiles = reinterpret_cast<double2*>(smem_raw);                          // NBUF * TILE amplitudes
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)NBUF * TILE * 16);  // full[NBUF], empty[NBUF]
    uint64_t* tbase_s = bars + 2 * NBUF;                                            // tile base per buffer
    DevRound* rounds_s = reinterpret_cast<DevRound*>(tbase_s + NBUF);               // nrounds

    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(&bars[b], 1);
            mbar_init(&bars[NBUF + b], NT / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(P.rounds);
        uint32_t* dst = reinterpret_cast<uint32_t*>(rounds_s);
        const int words = P.nrounds * (int)(sizeof(DevRound) / 4);
        for (int i = tid; i < words; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    if (tid >= NT) {
        // ---------------- producer warp: TMA bulk loads ----------------
        const int lane = tid - NT;
        uint32_t it = 0;
        for (uint64_t t = blockIdx.x; t < P.ntiles; t += gridDim.x, ++it) {
            const uint32_t buf = it % NBUF, use = it / NBUF;
            if (use > 0) mbar_wait(&bars[NBUF + buf], (use - 1) & 1);
            uint64_t base = 0;
            for (int s = 0; s < P.nseg; ++s) base |= ((t >> P.seg_src[s]) & P.seg_mask[s]) << P.seg_shift[s];
            if (lane == 0) {
                tbase_s[buf] = base;
                mbar_arrive_expect_tx(&bars[buf], (uint32_t)TILE * 16u);
            }
            __syncwarp();
            const uint32_t run_amps = P.run_bytes >> 4;
            for (int q = lane; q < P.nruns; q += 32)
                tma_bulk_g2s(tiles + (size_t)buf * TILE + (size_t)q * run_amps, P.state + base + __ldg(P.run_off + q),
                             P.run_bytes, &bars[buf]);
        }
        return;
    }

    // ---------------- consumers ----------------
    double2 a[R];
    uint32_t it = 0;
    for (uint64_t t = blockIdx.x; t < P.ntiles; t += gridDim.x, ++it) {
        const uint32_t buf = it % NBUF, use = it / NBUF;
        mbar_wait(&bars[buf], use & 1);
        const uint64_t tbase = tbase_s[buf];
        double2* sm = tiles + (size_t)buf * TILE;
        for (int r = 0; r < P.nrounds; ++r) {
            const DevRound& rd = rounds_s[r];
            const uint32_t tin = __ldg(P.tb + (size_t)(2 * r) * NT + tid);
            const uint64_t phys = tbase | __ldg(P.gt + (size_t)r * NT + tid);
#pragma unroll
            for (int i = 0; i < R; ++i) a[i] = sm[tin ^ rd.ro_in[i]];

            int op = rd.op_begin;
            const int op_end = rd.op_end;
            if (op < op_end) {
                DevOp cur = load_op(P.ops + op);
                for (; op < op_end; ++op) {
                    DevOp nxt;
                    if (op + 1 < op_end) nxt = load_op(P.ops + op + 1);
                    apply_op(a, cur, phys);
                    if (op + 1 < op_end) cur = nxt;
                }
            }

            if (rd.flags & 2u) {
                // last round: this thread is done with the buffer -> release it to the producer, then
                // write the 16 amplitudes back to HBM with 128-bit stores
                fence_proxy_async();
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(&bars[NBUF + buf]);
                double2* g = P.state + phys;
#pragma unroll
                for (int i = 0; i < R; ++i) g[rd.go[i]] = a[i];
            } else {
                const uint32_t tout = __ldg(P.tb + (size_t)(2 * r + 1) * NT + tid);
                if (rd.flags & 1u) consumer_sync<NT>();
#pragma unroll
                for (int i = 0; i < R; ++i) sm[tout ^ rd.ro_out[i]] = a[i];
                consumer_sync<NT>();
            }
        }
    }
}

// ---- host-side planning ----------------------------------------------------------------------------
struct HostGate {
    int target_phys, c1_phys, c2_phys;
    bool diag;
    double m[8];
    uint32_t kind;
};

static inline uint32_t swz(uint32_t j) { return j ^ (((j >> 3) ^ (j >> 6) ^ (j >> 9) ^ (j >> 12)) & 7u); }

static bool is_zero(double x) { return x == 0.0; }

// Pick the arithmetic class from the matrix itself (the type tag is only a hint).
static bool classify(const hq_gate& g, HostGate& h) {
    std::memcpy(h.m, g.mat, sizeof(h.m));
    const double* m = g.mat;
    const bool off0 = is_zero(m[2]) && is_zero(m[3]) && is_zero(m[4]) && is_zero(m[5]);
    const bool dia0 = is_zero(m[0]) && is_zero(m[1]) && is_zero(m[6]) && is_zero(m[7]);
    h.diag = off0 || g.target < 0;
    if (h.diag) {
        if (g.target < 0) {  // scalar: the reference keeps it in m00 (GCC, kernelOpt.cu:366)
            h.m[6] = h.m[0];
            h.m[7] = h.m[1];
        }
        const bool ident = h.m[0] == 1.0 && h.m[1] == 0.0 && h.m[6] == 1.0 && h.m[7] == 0.0;
        h.kind = OP_DIAG_R;
        return !ident;   // identity gates are dropped
    }
    const bool imag0 = is_zero(m[1]) && is_zero(m[3]) && is_zero(m[5]) && is_zero(m[7]);
    if (dia0 && m[2] == 1.0 && m[3] == 0.0 && m[4] == 1.0 && m[5] == 0.0) h.kind = OP_SWAP;
    else if (dia0 && m[2] == 0.0 && m[3] == -1.0 && m[4] == 0.0 && m[5] == 1.0) h.kind = OP_YL;
    else if (imag0) h.kind = OP_REAL;
    else if (is_zero(m[1]) && is_zero(m[2]) && is_zero(m[4]) && is_zero(m[7])) h.kind = OP_RXL;
    else h.kind = OP_GEN;
    return true;
}

}  // namespace hq

using namespace hq;

extern "C" int hq_group_tile_bits(void) { return rt().tile_bits; }
extern "C" int hq_group_min_run_bits(void) { return MIN_RUN_BITS; }

extern "C" int hq_group_plan_create(int L, uint64_t tile_mask, const hq_gate* gates, int ngates, hq_group_plan** out) {
    HQ_REQUIRE(out != nullptr, "plan out pointer is null");
    const int K = popcount64(tile_mask);
    HQ_REQUIRE(K >= 10 && K <= 12, "tile_mask must select 10, 11 or 12 bits");
    HQ_REQUIRE(L >= K && L <= 40, "local qubit count out of range for the gate-group kernel");
    
    HQ_REQUIRE((tile_mask >> L) == 0, "tile_mask has bits outside the local state");
    HQ_REQUIRE((tile_mask & ((1ull << MIN_RUN_BITS) - 1)) == ((1ull << MIN_RUN_BITS) - 1),
               "tile_mask must contain the low hq_group_min_run_bits() bits");
    HQ_REQUIRE(ngates >= 0 && (ngates == 0 || gates != nullptr), "bad gate list");
    const int NT = 1 << (K - RBITS);

    int phys_to_tile[64];
    for (int i = 0; i < 64; ++i) phys_to_tile[i] = -1;
    for (int b = 0, k = 0; b < L; ++b)
        if (tile_mask >> b & 1) phys_to_tile[b] = k++;

    // ---- classify + validate ----
    std::vector<HostGate> hg;
    hg.reserve(ngates);
    for (int i = 0; i < ngates; ++i) {
        const hq_gate& g = gates[i];
        HQ_REQUIRE(g.target >= -1 && g.target < L, "gate target outside the local state");
        HQ_REQUIRE(g.control >= -1 && g.control < L && g.control2 >= -1 && g.control2 < L, "gate control outside the local state");
        HostGate h{};
        h.target_phys = g.target; h.c1_phys = g.control; h.c2_phys = g.control2;
        if (!classify(g, h)) continue;
        HQ_REQUIRE(h.diag || phys_to_tile[g.target] >= 0, "non-diagonal gate target is not inside the tile");
        HQ_REQUIRE(g.target < 0 || (g.target != g.control && g.target != g.control2), "control equals target");
        hg.push_back(h);
    }

    // ---- split into rounds (<= RBITS distinct non-diagonal targets each, order-preserving up to commutation) ----
    struct Round { std::vector<int> reg; std::vector<int> gates; };
    std::vector<Round> rounds;
    {
        std::vector<int> remaining(hg.size());
        for (size_t i = 0; i < hg.size(); ++i) remaining[i] = (int)i;
        while (!remaining.empty()) {
            Round rd;
            uint64_t blockedX = 0, blockedZ = 0;
            std::vector<int> rest;
            for (int gi : remaining) {
                const HostGate& h = hg[gi];
                uint64_t q_nd = 0, q_d = 0;
                if (h.target_phys >= 0) (h.diag ? q_d : q_nd) |= 1ull << h.target_phys;
                if (h.c1_phys >= 0) q_d |= 1ull << h.c1_phys;
                if (h.c2_phys >= 0) q_d |= 1ull << h.c2_phys;
                bool can = !(q_nd & (blockedX | blockedZ)) && !(q_d & blockedX);
                if (can && !h.diag) {
                    const int tt = phys_to_tile[h.target_phys];
                    if (std::find(rd.reg.begin(), rd.reg.end(), tt) == rd.reg.end()) {
                        if ((int)rd.reg.size() < RBITS) rd.reg.push_back(tt);
                        else can = false;
                    }
                }
                if (can) rd.gates.push_back(gi);
                else { blockedX |= q_nd; blockedZ |= q_d; rest.push_back(gi); }
            }
            rounds.push_back(std::move(rd));
            remaining.swap(rest);
        }
        if (rounds.empty()) rounds.push_back(Round{});
        for (auto& rd : rounds) {   // pad the register set with the highest free tile bits
            for (int b = K - 1; b >= 0 && (int)rd.reg.size() < RBITS; --b)
                if (std::find(rd.reg.begin(), rd.reg.end(), b) == rd.reg.end()) rd.reg.push_back(b);
            std::sort(rd.reg.begin(), rd.reg.end());
        }
    }
    const int nrounds = (int)rounds.size();
    HQ_REQUIRE(nrounds <= 200, "too many rounds in one gate group");

    // ---- encode ----
    std::vector<DevRound> drounds(nrounds);
    std::vector<DevOp> dops;
    std::vector<uint16_t> tb((size_t)nrounds * 2 * NT);
    std::vector<uint64_t> gt((size_t)nrounds * NT);
    for (int r = 0; r < nrounds; ++r) {
        const Round& rd = rounds[r];
        int reg_of_tile[16];
        for (int i = 0; i < 16; ++i) reg_of_tile[i] = -1;
        for (int b = 0; b < RBITS; ++b) reg_of_tile[rd.reg[b]] = b;
        // thread-id bits -> tile bits: ascending, but the first three get distinct (bit mod 3) so that a
        // quarter-warp covers all eight 16-byte bank groups of the swizzled layout
        std::vector<int> tbits;
        for (int b = 0; b < K; ++b) if (reg_of_tile[b] < 0) tbits.push_back(b);
        {
            std::vector<int> first;
            bool used_res[3] = {false, false, false};
            for (int b : tbits) if (!used_res[b % 3] && (int)first.size() < 3) { first.push_back(b); used_res[b % 3] = true; }
            if (r == 0 && tbits.size() >= 3 && tbits[0] == 0 && tbits[1] == 1 && tbits[2] == 2) first = {0, 1, 2};
            std::vector<int> ordered = first;
            for (int b : tbits) if (std::find(first.begin(), first.end(), b) == first.end()) ordered.push_back(b);
            tbits.swap(ordered);
        }
        const bool lin_in = (r == 0);
        DevRound& d = drounds[r];
        std::memset(&d, 0, sizeof(d));
        for (int i = 0; i < R; ++i) {
            uint32_t j = 0;
            for (int b = 0; b < RBITS; ++b) if (i >> b & 1) j |= 1u << rd.reg[b];
            d.ro_in[i] = (uint16_t)(lin_in ? j : swz(j));
            d.ro_out[i] = (uint16_t)swz(j);
            d.go[i] = pdep64(j, tile_mask);
        }
        for (int t = 0; t < NT; ++t) {
            uint32_t j = 0;
            for (size_t b = 0; b < tbits.size(); ++b) if (t >> b & 1) j |= 1u << tbits[b];
            tb[(size_t)(2 * r) * NT + t] = (uint16_t)(lin_in ? j : swz(j));
            tb[(size_t)(2 * r + 1) * NT + t] = (uint16_t)swz(j);
            gt[(size_t)r * NT + t] = pdep64(j, tile_mask);
        }
        d.op_begin = (int)dops.size();
        for (int gi : rd.gates) {
            const HostGate& h = hg[gi];
            DevOp o{};
            std::memcpy(o.m, h.m, sizeof(o.m));
            o.kind = h.kind;
            for (int c : {h.c1_phys, h.c2_phys}) {
                if (c < 0) continue;
                const int tc = phys_to_tile[c];
                if (tc >= 0 && reg_of_tile[tc] >= 0) o.creg |= 1u << reg_of_tile[tc];
                else o.cphys |= 1ull << c;
            }
            if (h.diag) {
                if (h.m[0] == 1.0 && h.m[1] == 0.0) o.flags |= 1u;
                const int tt = h.target_phys >= 0 ? phys_to_tile[h.target_phys] : -1;
                if (tt >= 0 && reg_of_tile[tt] >= 0) { o.kind = OP_DIAG_R; o.tbit = reg_of_tile[tt]; }
                else { o.kind = OP_DIAG_T; o.tphys = h.target_phys >= 0 ? 1ull << h.target_phys : 0; }
            } else {
                o.tbit = reg_of_tile[phys_to_tile[h.target_phys]];
            }
            dops.push_back(o);
        }
        d.op_end = (int)dops.size();
        d.flags = (lin_in && nrounds > 1 ? 1u : 0u) | (r == nrounds - 1 ? 2u : 0u);
    }

    // ---- tile geometry ----
    int run_bits = 0;
    while (run_bits < K && (tile_mask >> run_bits & 1)) ++run_bits;
    const int nruns = 1 << (K - run_bits);
    std::vector<uint64_t> run_off(nruns);
    for (int q = 0; q < nruns; ++q) run_off[q] = pdep64((uint64_t)q << run_bits, tile_mask);

    auto* plan = new hq_group_plan();
    plan->L = L; plan->K = K; plan->NT = NT; plan->tile_mask = tile_mask;
    plan->nrounds = nrounds; plan->nops = (int)dops.size();
    GroupParams& p = plan->p;
    p.ntiles = 1ull << (L - K);
    p.nruns = nruns;
    p.run_bytes = 16u << run_bits;
    p.nrounds = nrounds;
    {   // tile number -> base: scatter over the runs of bits NOT in the tile
        const uint64_t outmask = ((L == 64 ? ~0ull : (1ull << L) - 1)) & ~tile_mask;
        int nseg = 0, src = 0, b = 0;
        while (b < L) {
            if (!(outmask >> b & 1)) { ++b; continue; }
            int e = b;
            while (e < L && (outmask >> e & 1)) ++e;
            p.seg_shift[nseg] = (uint8_t)b; p.seg_src[nseg] = (uint8_t)src; p.seg_mask[nseg] = (1ull << (e - b)) - 1;
            src += e - b; ++nseg; b = e;
        }
        p.nseg = nseg;
    }

    // one device blob: run_off | rounds | ops | gt | tb
    auto align16 = [](size_t x) { return (x + 15) & ~size_t(15); };
    const size_t o_run = 0;
    const size_t o_rounds = align16(o_run + run_off.size() * 8);
    const size_t o_ops = align16(o_rounds + drounds.size() * sizeof(DevRound));
    const size_t o_gt = align16(o_ops + std::max<size_t>(1, dops.size()) * sizeof(DevOp));
    const size_t o_tb = align16(o_gt + gt.size() * 8);
    const size_t total = align16(o_tb + tb.size() * 2);
    std::vector<unsigned char>& blob = plan->blob;
    blob.assign(total, 0);
    std::memcpy(blob.data() + o_run, run_off.data(), run_off.size() * 8);
    std::memcpy(blob.data() + o_rounds, drounds.data(), drounds.size() * sizeof(DevRound));
    if (!dops.empty()) std::memcpy(blob.data() + o_ops, dops.data(), dops.size() * sizeof(DevOp));
    std::memcpy(blob.data() + o_gt, gt.data(), gt.size() * 8);
    std::memcpy(blob.data() + o_tb, tb.data(), tb.size() * 2);
    plan->o_run = o_run; plan->o_rounds = o_rounds; plan->o_ops = o_ops; plan->o_gt = o_gt; plan->o_tb = o_tb;
    if (rt().ready) {   // without a bound GPU the plan is host-only (partitioner / planner tests)
        cudaError_t e = cudaMalloc(&plan->dev_blob, total);
        if (e == cudaSuccess) e = cudaMemcpyAsync(plan->dev_blob, blob.data(), total, cudaMemcpyHostToDevice, rt().compute);
        if (e == cudaSuccess) e = cudaStreamSynchronize(rt().compute);   // plan creation is off the hot path
        if (e != cudaSuccess) { delete plan; return cuda_fail(e, "plan upload", __FILE__, __LINE__); }
        unsigned char* d = static_cast<unsigned char*>(plan->dev_blob);
        p.run_off = reinterpret_cast<const uint64_t*>(d + o_run);
        p.rounds = reinterpret_cast<const DevRound*>(d + o_rounds);
        p.ops = reinterpret_cast<const DevOp*>(d + o_ops);
        p.gt = reinterpret_cast<const uint64_t*>(d + o_gt);
        p.tb = reinterpret_cast<const uint16_t*>(d + o_tb);
    }

    plan->smem = (size_t)NBUF * (16u << K) + (2 * NBUF + NBUF) * 8 + (size_t)nrounds * sizeof(DevRound) + 128;
    const uint64_t want = (uint64_t)(rt().ready ? rt().sm_count : 148) * (K <= 11 ? 2 : 1);
    plan->grid = (int)std::min<uint64_t>(p.ntiles, want);
    *out = plan;
    return HQ_OK;
}

template <int K>
static int launch_k(const hq_group_plan* plan, GroupParams p, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        HQ_CUDA(cudaFuncSetAttribute(group_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    group_kernel<K><<<plan->grid, plan->NT + 32, plan->smem, s>>>(p);
    HQ_CUDA(cudaGetLastError());
    return HQ_OK;
}

extern "C" int hq_group_plan_launch(const hq_group_plan* plan, void* state, int on_comm_stream) {
    HQ_REQUIRE(plan != nullptr && state != nullptr, "null plan or state");
    HQ_REQUIRE(rt().ready && plan->dev_blob != nullptr, "plan was created without a bound GPU (call hq_init first)");
    HQ_REQUIRE(plan->smem <= 227 * 1024, "gate group needs more shared memory than one SM has");
    GroupParams p = plan->p;
    p.state = static_cast<double2*>(state);
    cudaStream_t s = on_comm_stream ? rt().comm : rt().compute;
    switch (plan->K) {
        case 10: return launch_k<10>(plan, p, s);
        case 11: return launch_k<11>(plan, p, s);
        case 12: return launch_k<12>(plan, p, s);
        default: set_error("unsupported tile size"); return HQ_ERR_UNSUPPORTED;
    }
}

extern "C" int hq_group_plan_info(const hq_group_plan* plan, int* rounds, int* ops, int* grid, int* smem_bytes) {
    HQ_REQUIRE(plan != nullptr, "null plan");
    if (rounds) *rounds = plan->nrounds;
    if (ops) *ops = plan->nops;
    if (grid) *grid = plan->grid;
    if (smem_bytes) *smem_bytes = (int)plan->smem;
    return HQ_OK;
}

extern "C" int hq_group_plan_table_bytes(const hq_group_plan* plan, int* bytes) {
    HQ_REQUIRE(plan != nullptr && bytes != nullptr, "null argument");
    *bytes = (int)plan->blob.size();
    return HQ_OK;
}

extern "C" int hq_group_plan_destroy(hq_group_plan* plan) {
    if (!plan) return HQ_OK;
    if (plan->dev_blob) cudaFree(plan->dev_blob);
    delete plan;
    return HQ_OK;
}

extern "C" int hq_group_apply(void* state, int L, uint64_t tile_mask, const hq_gate* gates, int ngates) {
    hq_group_plan* plan = nullptr;
    int rc = hq_group_plan_create(L, tile_mask, gates, ngates, &plan);
    if (rc != HQ_OK) return rc;
    rc = hq_group_plan_launch(plan, state, 0);
    if (rc == HQ_OK) {
        cudaError_t e = cudaStreamSynchronize(rt().compute);
        if (e != cudaSuccess) rc = cuda_fail(e, "group sync", __FILE__, __LINE__);
    }
    hq_group_plan_destroy(plan);
    return rc;
}
