// Gate-group kernel for sm_100a (the OShareMem-class path).
//
// What it computes is what the reference's run<128> computes (src/kernelOpt.cu:388-433): for every
// tile of the local state (the amplitudes whose tile_mask bits vary while all other bits are fixed)
// apply, in order, a list of single-/controlled-qubit gates, in place.  How it does it is new:
//
//   * persistent CTAs (several per SM) loop over tiles; a tile is brought into shared memory by 1-D TMA
//     bulk copies (cp.async.bulk ... mbarrier::complete_tx, one per contiguous run of the tile), issued
//     as soon as the previous tile's last shared-memory read is done, so the HBM read of tile i+1 overlaps
//     the last round of arithmetic and the write-back of tile i (and the other CTAs of the SM);
//   * the host splits the gate list into ROUNDS.  In a round every thread holds 16 amplitudes
//     (4 "register qubits") in registers and applies all gates of the round there: no shared-memory
//     traffic and no barrier per gate (the reference does one SMEM read-modify-write plus a
//     __syncthreads() per gate, kernelOpt.cu:214-386).  Between rounds the tile is re-laid-out through
//     shared memory (XOR-swizzled, conflict-free 128-bit accesses) to change the register qubits;
//   * the lowered op list lives in shared memory; every (arithmetic class, target register bit, control
//     register bit) combination is its own straight-line body, all of them behind ONE brx.idx jump table, so a
//     gate costs its FP64 instructions plus a handful of decode instructions;
//   * gates of the form alpha*[[1,p],[q,-pq]] with p,q in {+-1} or {+-i} (H, RX/RY(+-pi/2)) are butterflies: two FP64
//     adds per amplitude and no multiply; the alphas are collected into one scalar per launch;
//   * diagonal gates that touch no register qubit of the round commute with everything else in it; they are
//     folded, per thread, into ONE complex factor (a "diagonal run") applied with a single multiply pass;
//   * controls and diagonal targets may sit on register bits, thread bits, or bits outside the tile
//     (the reference's "block bits"); they become predicate masks over the physical index;
//   * the last round writes registers straight back to HBM with 128-bit stores.
//
// Algorithmic traffic: 32 bytes per amplitude per launch (16 read + 16 written), independent of the
// number of gates.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <map>
#include <vector>

#include "hq_internal.h"
#include "group_plan.h"
#include "group_jit.h"

namespace hq {

// ---- PTX wrappers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
// volatile shared loads: issued exactly where written (the compiler otherwise sinks them behind the op-decode branches)
__device__ __forceinline__ void lds_v4(uint32_t a, uint4& v) {
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
}
__device__ __forceinline__ void lds_u32(uint32_t a, uint32_t& v) { asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); }
__device__ __forceinline__ void lds_v2u64(uint32_t a, ulonglong2& v) {
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(a));
}
__device__ __forceinline__ void lds_v2f64(uint32_t a, double2& v) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- in-register gate arithmetic -----------------------------------------------------------------
// The 16 amplitudes of a thread live in named PTX registers (hqa0..hqa31) declared once per kernel; all
// arithmetic on them is generated inline PTX (group_ops_gen.inc, produced by tools/gen_group_ops.py) that
// updates them IN PLACE, so the op loop carries no C++-visible state and a gate costs only its FP64 work.
//   real 2x2 [[a,b],[c,d]] on scalars (u,v):  t <- u ; u <- a*u + b*v ; v <- c*t + d*v   (4 FP64 + one register copy)
}  // namespace hq
#if HQ_RBITS == 4
#include "group_ops_gen_r4.inc"
#else
#include "group_ops_gen_r3.inc"
#endif
namespace hq {

// A run of diagonal gates none of which touches a register bit: every amplitude of the thread gets the same factor.
// With creg != 0 the factor applies only to the amplitudes whose register-index bits creg are all 1: every diagonal gate
// with one operand on a register bit and the others elsewhere (the cu1 ladder of a QFT, CZ / CRZ fans) joins such a run.
__device__ __forceinline__ void op_diag_run(uint32_t entries_sa, int n, uint64_t phys, uint32_t creg) {
    double fr = 1.0, fi = 0.0;
    bool any = false;
    for (int e = 0; e < n; ++e, entries_sa += (uint32_t)sizeof(DevOp)) {   // entries are read by shared address: no generic pointer kept live
        ulonglong2 cp;      // cphys, tphys
        lds_v2u64(entries_sa + 80, cp);
        if ((phys & cp.x) != cp.x) continue;
        const bool hi = (cp.y == 0) || (phys & cp.y);
        uint32_t flags;
        lds_u32(entries_sa + 72, flags);
        if (!hi && (flags & 1u)) continue;
        double2 d;
        lds_v2f64(entries_sa + (hi ? 48u : 0u), d);
        const double nr = fma(-fi, d.y, fr * d.x);
        fi = fma(fi, d.x, fr * d.y);
        fr = nr;
        any = true;
    }
    if (!any) return;
    if (creg == 0) hq_cmul_all(fr, fi);
    else if ((creg & (creg - 1)) == 0) hq_cmul_bit(fr, fi, 31 - __clz(creg));
    else hq_cmul_masked(fr, fi, creg);
}

// ---- the kernel ----------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void issue_tile_load(const GroupParams& P, uint64_t t, double2* tile, uint64_t* bar,
                                                uint64_t* tbase_s, int lane) {
    uint64_t base = P.fixed_base;
    for (int s = 0; s < P.nseg; ++s) base |= ((t >> P.seg_src[s]) & P.seg_mask[s]) << P.seg_shift[s];
    if (lane == 0) {
        *tbase_s = base;
        mbar_arrive_expect_tx(bar, (16u << K));
    }
    __syncwarp();
    const uint32_t run_amps = P.run_bytes >> 4;
    for (int q = lane; q < P.nruns; q += 32)
        tma_bulk_g2s(tile + (size_t)q * run_amps, P.state + base + __ldg(P.run_off + q), P.run_bytes, bar);
}

template <int K, int MINB>
__global__ void __launch_bounds__(1 << (K - RBITS), MINB) group_kernel(const __grid_constant__ GroupParams P) {
    constexpr int NT = 1 << (K - RBITS);
    constexpr int TILE = 1 << K;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2* tile = reinterpret_cast<double2*>(smem_raw);                              // TILE amplitudes
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)TILE * 16);        // "tile landed" mbarrier
    uint64_t* tbase_s = bar + 1;                                                       // base index of the landed tile
    DevRound* rounds_s = reinterpret_cast<DevRound*>(bar + 2);
    DevOp* ops_s = reinterpret_cast<DevOp*>(rounds_s + P.nrounds);
    // per-round, per-thread index tables (the same for every tile): kept in shared memory so that a round starts after a
    // shared-memory latency, not an L2 one (the tile write-back streams through L1 and evicts anything cached there)
    uint64_t* gt_s = reinterpret_cast<uint64_t*>(ops_s + P.nops + 1);                  // +1: read-ahead slot
    uint16_t* tb_s = reinterpret_cast<uint16_t*>(gt_s + (size_t)P.nrounds * NT);

    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // stage the round descriptors and the op list once per CTA
        const uint4* src = reinterpret_cast<const uint4*>(P.rounds);
        uint4* dst = reinterpret_cast<uint4*>(rounds_s);
        const int n16 = P.nrounds * (int)(sizeof(DevRound) / 16);
        for (int i = tid; i < n16; i += NT) dst[i] = __ldg(src + i);
        src = reinterpret_cast<const uint4*>(P.ops);
        dst = reinterpret_cast<uint4*>(ops_s);
        const int m16 = P.nops * (int)(sizeof(DevOp) / 16);
        for (int i = tid; i < m16; i += NT) dst[i] = __ldg(src + i);
        for (int r = 0; r < P.nrounds; ++r) {
            gt_s[(size_t)r * NT + tid] = __ldg(P.gt + (size_t)r * NT + tid);
            tb_s[(size_t)(2 * r) * NT + tid] = __ldg(P.tb + (size_t)(2 * r) * NT + tid);
            tb_s[(size_t)(2 * r + 1) * NT + tid] = __ldg(P.tb + (size_t)(2 * r + 1) * NT + tid);
        }
    }
    __syncthreads();

    uint64_t t = blockIdx.x;
    if (t < P.ntiles && tid < 32) issue_tile_load<K>(P, t, tile, bar, tbase_s, tid);

    HQ_DECLARE_AMP_REGS();
    const uint32_t tile_s = smem_u32(tile);
    const uint32_t ops_sa = smem_u32(ops_s);
    for (uint32_t it = 0; t < P.ntiles; t += gridDim.x, ++it) {
        mbar_wait(bar, it & 1);
        const uint64_t tbase = *tbase_s;
        for (int r = 0; r < P.nrounds; ++r) {
            const DevRound& rd = rounds_s[r];
            const uint32_t tin = tb_s[(size_t)(2 * r) * NT + tid];
            const uint64_t phys = tbase | gt_s[(size_t)r * NT + tid];
            hq_load_amps(tile_s, tin, rd.ro_in);

            const bool last = rd.flags & 2u;
            if (last) {
                // every thread has its amplitudes in registers: the buffer can take the next tile now, so its HBM
                // read runs under this round's arithmetic and write-back
                fence_proxy_async();
                __syncthreads();
                const uint64_t tn = t + gridDim.x;
                if (tn < P.ntiles && tid < 32) issue_tile_load<K>(P, tn, tile, bar, tbase_s, tid);
            }

            // The op loop.  A plain op (register target, no predicate outside the registers: most single-qubit gates) costs
            // one header read-ahead, one flag test and the indexed jump into its body; the bodies fetch their own
            // coefficients.  Everything else -- predicate masks, diagonal gates on non-register bits, diagonal runs --
            // hides behind ONE "special" flag bit.  The header is read one op AHEAD, while the current body runs.
            uint32_t pa = ops_sa + (uint32_t)rd.op_begin * (uint32_t)sizeof(DevOp);
            const uint32_t pa_end = ops_sa + (uint32_t)rd.op_end * (uint32_t)sizeof(DevOp);
            uint4 nhd;         // code, creg, flags, aux
            lds_v4(pa + 64, nhd);
            while (pa < pa_end) {
                const uint4 hd = nhd;
                const uint32_t cur = pa;
                pa += (uint32_t)sizeof(DevOp);
                lds_v4(pa + 64, nhd);   // (the slot after the last op is readable padding, never interpreted)
                if (hd.z & 4u) {
                    if (hd.x == CODE_DIAG_RUN) {
                        uint32_t n;   // entry count: re-read here rather than kept live through every plain op (a spill otherwise)
                        lds_u32(cur + 76, n);
                        pa += n * (uint32_t)sizeof(DevOp);
                        lds_v4(pa + 64, nhd);   // the header read ahead above was the run's first entry: fetch the real next op
                        op_diag_run(cur + (uint32_t)sizeof(DevOp), (int)n, phys, hd.y);
                        continue;
                    }
                    // predicate masks (cphys, tphys) are not read ahead: four more registers carried around the loop push
                    // ptxas into spilling at the 128-register cap
                    ulonglong2 cp;
                    lds_v2u64(cur + 80, cp);
                    if ((phys & cp.x) != cp.x) continue;
                    if (hd.x == CODE_DIAG_T) {
                        const bool hi = (cp.y == 0) || (phys & cp.y);
                        if (!hi && (hd.z & 1u)) continue;
                        double2 d;
                        lds_v2f64(cur + (hi ? 48u : 0u), d);
                        hq_cmul_masked(d.x, d.y, hd.y);
                        continue;
                    }
                }
                hq_apply_op(hd.z >> 16, hd.y, cur);
            }

            if (last) {
                hq_store_amps_global(P.state + phys, rd.go);
            } else {
                const uint32_t tout = tb_s[(size_t)(2 * r + 1) * NT + tid];
                // Exchanges whose data stays inside each warp's own slice of the tile (same warp qubits before and after:
                // flags bits 2/3, proven by the planner) need no CTA barrier: the warps of a CTA drift apart and fill each
                // other's shared-memory and barrier bubbles with FP64 work.
                if (rd.flags & 1u) { if (rd.flags & 8u) __syncwarp(); else __syncthreads(); }
                hq_store_amps(tile_s, tout, rd.ro_out);
                if (rd.flags & 4u) __syncwarp(); else __syncthreads();
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
    double alpha[2];   // butterflies: the scalar taken out of the matrix (deferred to one factor per launch)
};

static inline uint32_t swz(uint32_t j) { return j ^ (((j >> 3) ^ (j >> 6) ^ (j >> 9) ^ (j >> 12)) & 7u); }

static bool is_zero(double x) { return x == 0.0; }
static bool rt_no_butterfly() { static const bool off = getenv("HQ_NO_BUTTERFLY") != nullptr; return off; }

// Pick the arithmetic class from the matrix itself (the type tag is only a hint).
static bool classify(const hq_gate& g, HostGate& h, std::complex<double>& launch_scale, bool& any_scale) {
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
        // An uncontrolled diag(d0, d1) = d0 * diag(1, d1/d0): the scalar joins the launch's one deferred factor and what is left
        // multiplies only the "hi" half -- and can merge into the per-thread / per-register-bit diagonal runs (RZ, the first
        // factor of a rewritten ZZ rotation, global phases: all free of FP64 work on the "lo" half from here on).
        if (g.control < 0 && g.control2 < 0 && !(h.m[0] == 1.0 && h.m[1] == 0.0) && std::hypot(h.m[0], h.m[1]) > 0.5) {
            const std::complex<double> d0(h.m[0], h.m[1]), r = std::complex<double>(h.m[6], h.m[7]) / d0;
            launch_scale *= d0;
            any_scale = true;
            h.m[0] = 1.0; h.m[1] = 0.0;
            h.m[6] = r.real(); h.m[7] = r.imag();
            if (std::abs(r - 1.0) < 1e-15) { h.m[6] = 1.0; h.m[7] = 0.0; }   // a pure scalar gate: nothing left to apply
        }
        const bool ident = h.m[0] == 1.0 && h.m[1] == 0.0 && h.m[6] == 1.0 && h.m[7] == 0.0;
        h.kind = OP_DIAG_R;
        return !ident;   // identity gates are dropped
    }
    const bool imag0 = is_zero(m[1]) && is_zero(m[3]) && is_zero(m[5]) && is_zero(m[7]);
    if (g.control < 0 && g.control2 < 0 && !rt_no_butterfly()) {
        // alpha * [[1, p],[q, -p q]] with p, q both real units or both imaginary units?  (tolerance: cos(pi/4) and
        // sin(pi/4) differ in the last bit; 1e-14 relative is far inside the 1e-10 amplitude budget)
        typedef std::complex<double> C;
        const C a(m[0], m[1]), b(m[2], m[3]), c(m[4], m[5]), d(m[6], m[7]);
        if (std::abs(a) > 0.5) {
            const C pp = b / a, qq = c / a, rr = d / a;
            for (int v = 0; v < 8; ++v) {
                const C pv(HQ_BF_PQ[v][0], HQ_BF_PQ[v][1]), qv(HQ_BF_PQ[v][2], HQ_BF_PQ[v][3]);
                if (std::abs(pp - pv) < 1e-14 && std::abs(qq - qv) < 1e-14 && std::abs(rr + pv * qv) < 1e-14) {
                    h.kind = OP_BF0 + v;
                    h.alpha[0] = a.real(); h.alpha[1] = a.imag();
                    return true;
                }
            }
        }
    }
    if (dia0 && m[2] == 1.0 && m[3] == 0.0 && m[4] == 1.0 && m[5] == 0.0) h.kind = OP_SWAP;
    else if (dia0 && m[2] == 0.0 && m[3] == -1.0 && m[4] == 0.0 && m[5] == 1.0) h.kind = OP_YL;
    else if (imag0) h.kind = OP_REAL;
    else if (is_zero(m[1]) && is_zero(m[2]) && is_zero(m[4]) && is_zero(m[7])) h.kind = OP_RXL;
    else h.kind = OP_GEN;
    return true;
}

// Coefficients of the 2x2 bodies: OP_REAL {a, b, c, d} = the real matrix; OP_RXL {a, b, c, d} of [[a, i b], [i c, d]];
// OP_GEN the complex matrix itself (row-major re, im).
static void encode_2x2(uint32_t kind, const double* M, double* out) {
    std::memset(out, 0, 8 * sizeof(double));
    if (kind == OP_REAL) { out[0] = M[0]; out[1] = M[2]; out[2] = M[4]; out[3] = M[6]; }
    else if (kind == OP_RXL) { out[0] = M[0]; out[1] = M[3]; out[2] = M[5]; out[3] = M[6]; }
    else std::memcpy(out, M, 8 * sizeof(double));
}

}  // namespace hq

using namespace hq;

static std::string plan_identity(const hq_group_plan& plan);

extern "C" int hq_group_tile_bits(void) { return rt().tile_bits; }
extern "C" int hq_group_min_run_bits(void) { return MIN_RUN_BITS; }

extern "C" int hq_group_plan_create(int L, uint64_t tile_mask, const hq_gate* gates, int ngates, hq_group_plan** out) {
    return hq_group_plan_create_ex(L, tile_mask, 0, 0, gates, ngates, out);
}

// fixed_mask / fixed_value: physical local bits that do NOT vary in this launch (the launch covers only the amplitudes
// whose fixed bits equal fixed_value): one chunk of a state whose other chunks are still being exchanged.
extern "C" int hq_group_plan_create_ex(int L, uint64_t tile_mask, uint64_t fixed_mask, uint64_t fixed_value, const hq_gate* gates,
                                       int ngates, hq_group_plan** out) {
    HQ_REQUIRE(out != nullptr, "plan out pointer is null");
    HQ_REQUIRE((fixed_mask >> L) == 0 && (fixed_mask & tile_mask) == 0 && (fixed_value & ~fixed_mask) == 0,
               "fixed bits must be local, outside the tile, and fixed_value within fixed_mask");
    const int K = popcount64(tile_mask);
    HQ_REQUIRE(K >= 10 && K <= 12, "tile_mask must select 10, 11 or 12 bits");
    HQ_REQUIRE(L - popcount64(fixed_mask) >= K && L <= 40, "local qubit count out of range for the gate-group kernel");
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
    std::complex<double> launch_scale(1.0, 0.0);   // scalars taken out of gates (diagonals here, butterflies below)
    bool any_bfly = false;                          // "the launch has a deferred scalar"
    std::vector<HostGate> hg;
    hg.reserve(ngates);
    for (int i = 0; i < ngates; ++i) {
        const hq_gate& g = gates[i];
        HQ_REQUIRE(g.target >= -1 && g.target < L, "gate target outside the local state");
        HQ_REQUIRE(g.control >= -1 && g.control < L && g.control2 >= -1 && g.control2 < L, "gate control outside the local state");
        for (int q : {g.target, g.control, g.control2})
            HQ_REQUIRE(q < 0 || !(fixed_mask >> q & 1), "gate touches a fixed bit (resolve it when lowering the gate)");
        HostGate h{};
        h.target_phys = g.target; h.c1_phys = g.control; h.c2_phys = g.control2;
        if (!classify(g, h, launch_scale, any_bfly)) continue;
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
        const bool avoid_low_in_round0 = getenv("HQ_ROUND0_LOW_BITS") == nullptr;
        const bool pack_rounds = getenv("HQ_NO_ROUND_PACK") == nullptr;
        // One pass over the remaining gates: a gate joins the round when nothing it fails to commute with was left behind and,
        // if it is non-diagonal, its target is (or can still become) one of the round's <= RBITS register qubits.  `allowed`
        // restricts which tile bits may become register qubits (all ones = first come, first served).
        auto fill = [&](const std::vector<int>& remaining, uint32_t allowed, bool first_round, Round* out, std::vector<int>* rest) {
            uint64_t blockedX = 0, blockedZ = 0;
            std::vector<int> reg;
            int taken = 0;
            for (int gi : remaining) {
                const HostGate& h = hg[gi];
                uint64_t q_nd = 0, q_d = 0;
                if (h.target_phys >= 0) (h.diag ? q_d : q_nd) |= 1ull << h.target_phys;
                if (h.c1_phys >= 0) q_d |= 1ull << h.c1_phys;
                if (h.c2_phys >= 0) q_d |= 1ull << h.c2_phys;
                bool can = !(q_nd & (blockedX | blockedZ)) && !(q_d & blockedX);
                if (can && !h.diag) {
                    const int tt = phys_to_tile[h.target_phys];
                    // Round 0 reads the linear TMA image: a register qubit on tile bits 0..2 would put every lane of a
                    // quarter-warp on the same 16-byte bank group (8-way conflicts on all 16 loads; ncu: 3x the ideal
                    // wavefronts over a 4-round launch).  Those gates wait for round 1, which reads the swizzled layout.
                    if (first_round && tt < 3 && avoid_low_in_round0) can = false;
                    if (can && !(allowed >> tt & 1)) can = false;
                    if (can && std::find(reg.begin(), reg.end(), tt) == reg.end()) {
                        if ((int)reg.size() < RBITS) reg.push_back(tt);
                        else can = false;
                    }
                }
                if (can) { ++taken; if (out) out->gates.push_back(gi); }
                else { blockedX |= q_nd; blockedZ |= q_d; if (rest) rest->push_back(gi); }
            }
            if (out) out->reg = reg;
            return taken;
        };
        while (!remaining.empty()) {
            const bool first_round = rounds.empty();
            uint32_t allowed = ~0u;
            if (pack_rounds) {
                // Which register qubits?  First come, first served often strands gates behind a qubit that did not make
                // the cut.  Grow the set one qubit at a time, each time taking the qubit that lets the round hold the most
                // gates; keep the result when it beats first-come.
                uint32_t targets = 0;
                for (int gi : remaining) if (!hg[gi].diag) targets |= 1u << phys_to_tile[hg[gi].target_phys];
                if (popcount64(targets) > RBITS) {
                    uint32_t set = 0;
                    int best_total = fill(remaining, 0u, first_round, nullptr, nullptr);
                    for (int step = 0; step < RBITS; ++step) {
                        int best_bit = -1, best_cnt = best_total;
                        for (int b = 0; b < K; ++b) {
                            if (!(targets >> b & 1) || (set >> b & 1)) continue;
                            const int cnt = fill(remaining, set | 1u << b, first_round, nullptr, nullptr);
                            if (cnt > best_cnt) { best_cnt = cnt; best_bit = b; }
                        }
                        if (best_bit < 0) break;
                        set |= 1u << best_bit;
                        best_total = best_cnt;
                    }
                    if (best_total > fill(remaining, ~0u, first_round, nullptr, nullptr)) allowed = set;
                }
            }
            Round rd;
            std::vector<int> rest;
            fill(remaining, allowed, first_round, &rd, &rest);
            rounds.push_back(std::move(rd));
            remaining.swap(rest);
        }
        if (rounds.empty()) rounds.push_back(Round{});
    }
    const int nrounds = (int)rounds.size();
    // Warp qubits (thread-id bits 5 and up) of every round.  A round whose warp qubits equal the previous round's exchanges
    // data only inside each warp, so the barrier between them is a __syncwarp.  Keep them while none is needed as a register
    // qubit; when they must change, take the candidates that stay clear of the following rounds' targets the longest (ties:
    // the highest, so that lanes keep the low bits and HBM stores stay contiguous).  Tile bits 0..2 never serve: the swizzle
    // only rewrites bits 0..2, so avoiding them keeps a warp's slice the same set of positions in the linear (TMA) image and
    // in the swizzled one.
    const int nwb = K - RBITS - 5;
    const bool allow_local = nwb > 0 && getenv("HQ_NO_LOCAL_EXCHANGE") == nullptr;
    std::vector<std::vector<int>> warp_bits(nrounds);
    std::vector<char> local_in(nrounds, 0);   // round r reads only what the same warp wrote in round r-1
    if (allow_local) {
        auto needs = [&](int r, int b) { return std::find(rounds[r].reg.begin(), rounds[r].reg.end(), b) != rounds[r].reg.end(); };
        for (int r = 0; r < nrounds; ++r) {
            bool sticky = r > 0;
            if (sticky) for (int b : warp_bits[r - 1]) if (needs(r, b)) sticky = false;
            if (sticky) { warp_bits[r] = warp_bits[r - 1]; local_in[r] = 1; continue; }
            std::vector<std::pair<int, int>> cand;   // (-survival, -bit)
            for (int b = 3; b < K; ++b) {
                if (needs(r, b)) continue;
                int surv = 0;
                for (int rr = r + 1; rr < nrounds && !needs(rr, b); ++rr) ++surv;
                cand.push_back({-surv, -b});
            }
            std::sort(cand.begin(), cand.end());
            for (int i = 0; i < nwb; ++i) warp_bits[r].push_back(-cand[i].second);   // K - 3 - RBITS >= nwb candidates always exist
            std::sort(warp_bits[r].begin(), warp_bits[r].end());
        }
    }
    for (int r = 0; r < nrounds; ++r) {   // pad the register set with the highest tile bits that are not warp qubits
        Round& rd = rounds[r];
        for (int b = K - 1; b >= 0 && (int)rd.reg.size() < RBITS; --b)
            if (std::find(rd.reg.begin(), rd.reg.end(), b) == rd.reg.end() &&
                std::find(warp_bits[r].begin(), warp_bits[r].end(), b) == warp_bits[r].end())
                rd.reg.push_back(b);
        std::sort(rd.reg.begin(), rd.reg.end());
    }
    // butterflies leave their scalars behind too: one factor for the whole launch, applied with the last round's diagonal run
    for (const HostGate& h : hg)
        if (h.kind >= OP_BF0 && h.kind <= OP_BF7 && !h.diag) { launch_scale *= std::complex<double>(h.alpha[0], h.alpha[1]); any_bfly = true; }
    const bool trace_plan = getenv("HQ_TRACE_PLAN") != nullptr;

    // ---- encode ----
    std::vector<DevRound> drounds(nrounds);
    std::vector<DevOp> dops;
    std::vector<uint16_t> tb((size_t)nrounds * 2 * NT);
    std::vector<uint64_t> gt((size_t)nrounds * NT);
    std::vector<hq_group_plan::RoundMeta> meta(nrounds);
    for (int r = 0; r < nrounds; ++r) {
        const Round& rd = rounds[r];
        int reg_of_tile[16];
        for (int i = 0; i < 16; ++i) reg_of_tile[i] = -1;
        for (int b = 0; b < RBITS; ++b) reg_of_tile[rd.reg[b]] = b;

        // lower the round's gates first: the thread-bit assignment below wants to know which tile bits act as
        // thread-level predicates
        std::vector<DevOp> body, run;
        // creg -> diagonal gates waiting to be merged into one run: (run entry, the op as it would be emitted on its own)
        std::map<uint32_t, std::vector<std::pair<DevOp, DevOp>>> pending;
        auto flush = [&](uint32_t touching) {             // emit every pending run that involves register bits `touching`
            for (auto it = pending.begin(); it != pending.end();) {
                if (!(it->first & touching)) { ++it; continue; }
                if (it->second.size() == 1) {
                    body.push_back(it->second[0].second);
                } else {
                    DevOp hdr{};
                    hdr.code = CODE_DIAG_RUN;
                    hdr.aux = (uint32_t)it->second.size();
                    hdr.creg = it->first;
                    body.push_back(hdr);
                    for (auto& pr : it->second) body.push_back(pr.first);
                }
                it = pending.erase(it);
            }
        };
        uint32_t predicate_tile_bits = 0;
        for (int gi : rd.gates) {
            const HostGate& h = hg[gi];
            DevOp o{};
            std::memcpy(o.m, h.m, sizeof(o.m));
            // A controlled diag(1,d) is symmetric in its qubits, so any of them may play "target": prefer one that
            // is a register bit (then the others are plain predicates).
            int tgt = h.target_phys, ctl[2] = {h.c1_phys, h.c2_phys};
            const bool d0one = h.diag && h.m[0] == 1.0 && h.m[1] == 0.0;
            auto in_reg = [&](int phys) { return phys >= 0 && phys_to_tile[phys] >= 0 && reg_of_tile[phys_to_tile[phys]] >= 0; };
            if (h.diag && d0one && tgt >= 0 && !in_reg(tgt))
                for (int& c : ctl) if (in_reg(c)) { std::swap(tgt, c); break; }
            for (int c : ctl) {
                if (c < 0) continue;
                if (in_reg(c)) o.creg |= 1u << reg_of_tile[phys_to_tile[c]];
                else { o.cphys |= 1ull << c; if (phys_to_tile[c] >= 0) predicate_tile_bits |= 1u << phys_to_tile[c]; }
            }
            const int ncreg = popcount64(o.creg);
            const uint32_t cbc = ncreg == 0 ? 0 : (ncreg == 1 ? 1 + (uint32_t)__builtin_ctz(o.creg) : CBC_GENERIC);
            if (h.diag) {
                if (d0one) o.flags |= 1u;
                if (in_reg(tgt)) {
                    const uint32_t tbit = reg_of_tile[phys_to_tile[tgt]];
                    const bool zflip = d0one && h.m[6] == -1.0 && h.m[7] == 0.0;
                    o.code = op_code(zflip ? OP_ZFLIP : (d0one ? OP_DIAG_R1 : OP_DIAG_R), tbit, cbc);
                    if (d0one && o.creg == 0) {
                        // diag(1,d) whose ONLY register operand is tbit (T on a register qubit, the cu1 ladder of a QFT,
                        // CZ fans): "multiply by d where tbit = 1 and the other operands are 1".  Consecutive ones on the
                        // same register bit merge into a single run: one factor per thread, one multiply pass.
                        DevOp e = o;
                        e.code = CODE_DIAG_T;
                        e.creg = 1u << tbit;
                        e.tphys = 0;   // scalar d1, gated by cphys
                        pending[e.creg].push_back({e, o});
                    } else {
                        body.push_back(o);
                    }
                } else {
                    o.tphys = tgt >= 0 ? 1ull << tgt : 0;
                    if (tgt >= 0 && phys_to_tile[tgt] >= 0) predicate_tile_bits |= 1u << phys_to_tile[tgt];
                    o.code = CODE_DIAG_T;
                    // no register bit involved: joins the round's diagonal run; register bits only as controls: joins the
                    // pending run of that control set (flushed before the next non-diagonal gate on one of those bits)
                    if (o.creg == 0) run.push_back(o);
                    else pending[o.creg].push_back({o, o});
                }
            } else {
                const uint32_t tbit = reg_of_tile[phys_to_tile[tgt]];
                flush(1u << tbit);   // diagonal runs controlled by this register bit must act before it is mixed
                const bool generic_only = h.kind == OP_GEN || h.kind == OP_REAL || h.kind == OP_RXL || h.kind == OP_YL;
                const uint32_t cb = (generic_only && cbc != 0) ? CBC_GENERIC : cbc;
                if (h.kind == OP_SWAP || h.kind == OP_YL) {
                    o.code = op_code(h.kind, tbit, cb);
                    body.push_back(o);
                } else if (h.kind >= OP_BF0 && h.kind <= OP_BF7) {
                    std::memset(o.m, 0, sizeof(o.m));
                    o.code = op_code(h.kind, tbit, 0);
                    body.push_back(o);
                } else {
                    encode_2x2(h.kind, h.m, o.m);
                    o.code = op_code(h.kind, tbit, cb);
                    body.push_back(o);
                }
            }
        }

        // thread-id bits -> tile bits.  Constraints: the first three lane bits get distinct (bit mod 3) so that a
        // quarter-warp covers all eight 16-byte bank groups of the swizzled layout; round 0 reads the linear TMA
        // image, so it wants tile bits 0,1,2 on lanes 0..2; the last round stores to HBM, so its lanes are the
        // lowest tile bits (contiguous 512-byte warp stores).  Otherwise prefer bits that are NOT predicates on the
        // lanes, so that controlled gates switch whole warps on/off instead of diverging inside a warp.
        std::vector<int> free_bits;
        for (int b = 0; b < K; ++b) if (reg_of_tile[b] < 0) free_bits.push_back(b);
        std::vector<int> tbits;
        const bool is_last = r == nrounds - 1, lin_in = r == 0;
        if (allow_local) {
            std::vector<int> lanes_free;
            for (int b : free_bits) if (std::find(warp_bits[r].begin(), warp_bits[r].end(), b) == warp_bits[r].end()) lanes_free.push_back(b);
            free_bits.swap(lanes_free);   // the rules below now order the LANE bits only; the warp bits are appended after them
        }
        if (is_last) {
            tbits = free_bits;
        } else {
            std::vector<int> pref = free_bits;
            std::stable_sort(pref.begin(), pref.end(), [&](int x, int y) {
                return (predicate_tile_bits >> x & 1) < (predicate_tile_bits >> y & 1);
            });
            std::vector<int> first;
            bool used_res[3] = {false, false, false};
            if (lin_in) { for (int b : free_bits) if (b < 3) { first.push_back(b); used_res[b % 3] = true; } }
            for (int b : pref) if ((int)first.size() < 3 && !used_res[b % 3]) { first.push_back(b); used_res[b % 3] = true; }
            tbits = first;
            for (int b : pref) if (std::find(tbits.begin(), tbits.end(), b) == tbits.end()) tbits.push_back(b);
        }
        if (allow_local) tbits.insert(tbits.end(), warp_bits[r].begin(), warp_bits[r].end());

        meta[r].tbits = tbits;
        for (int b = 0; b < RBITS; ++b) meta[r].reg[b] = rd.reg[b];
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
        flush(~0u);
        if (is_last && any_bfly) {   // scalar entry (tphys = 0: always d1, no predicate) of the round's diagonal run
            DevOp sc{};
            sc.code = CODE_DIAG_T;
            sc.m[0] = sc.m[6] = launch_scale.real();
            sc.m[1] = sc.m[7] = launch_scale.imag();
            run.push_back(sc);
        }
        for (DevOp& o : body) {   // body index for the kernel's single indexed branch; "special" bit (see the op loop)
            if (o.code < CODE_DIAG_T) o.flags |= (uint32_t)HQ_OP_BODY_INDEX[o.code] << 16;
            if ((o.cphys | o.tphys) || o.code >= CODE_DIAG_T) o.flags |= 4u;
        }
        d.op_begin = (int)dops.size();
        if (!run.empty()) {   // the run commutes with every other op of the round (it touches no register bit)
            DevOp hdr{};
            hdr.code = CODE_DIAG_RUN;
            hdr.aux = (uint32_t)run.size();
            hdr.flags = 4u;
            dops.push_back(hdr);
            dops.insert(dops.end(), run.begin(), run.end());
        }
        dops.insert(dops.end(), body.begin(), body.end());
        d.op_end = (int)dops.size();
        if (trace_plan) {   // developer aid: which bodies does this round execute?
            fprintf(stderr, "[plan] round %d/%d%s regs={", r, nrounds, local_in[r] ? " (warp-local in)" : "");
            for (int b = 0; b < RBITS; ++b) fprintf(stderr, "%d%s", rd.reg[b], b + 1 < RBITS ? "," : "} ops:");
            for (int i = d.op_begin; i < d.op_end; ++i) fprintf(stderr, " %u", dops[i].code);
            fprintf(stderr, "\n");
        }
        d.flags = (lin_in && nrounds > 1 ? 1u : 0u) | (is_last ? 2u : 0u);
    }

    // ---- warp-local exchanges: prove them on the tables just built, then flag them ----
    // Exchange r -> r+1 is warp-local iff, for every warp, the shared-memory positions it writes at the end of round r are
    // exactly the positions it reads at the start of round r+1.  Round 0's layout change (linear -> swizzled) needs, in
    // addition, that the warp reads and writes the same positions within round 0.
    int nlocal = 0;
    {
        auto positions = [&](int r, int which, int warp) {   // which: 0 = read layout, 1 = write layout
            std::vector<uint16_t> v;
            v.reserve(32 * R);
            const uint16_t* ro = which ? drounds[r].ro_out : drounds[r].ro_in;
            for (int t = warp * 32; t < warp * 32 + 32 && t < NT; ++t)
                for (int i = 0; i < R; ++i) v.push_back((uint16_t)(tb[(size_t)(2 * r + which) * NT + t] ^ ro[i]));
            std::sort(v.begin(), v.end());
            return v;
        };
        const int nwarps = (NT + 31) / 32;
        for (int r = 0; r + 1 < nrounds; ++r) {
            if (!local_in[r + 1]) continue;
            bool ok = true;
            for (int w = 0; w < nwarps && ok; ++w) ok = positions(r, 1, w) == positions(r + 1, 0, w);
            if (!ok) continue;
            drounds[r].flags |= 4u;
            ++nlocal;
            if (drounds[r].flags & 1u) {
                bool same = true;
                for (int w = 0; w < nwarps && same; ++w) same = positions(r, 0, w) == positions(r, 1, w);
                if (same) drounds[r].flags |= 8u;
            }
        }
    }

    // ---- tile geometry ----
    int run_bits = 0;
    while (run_bits < K && (tile_mask >> run_bits & 1)) ++run_bits;
    const int nruns = 1 << (K - run_bits);
    std::vector<uint64_t> run_off(nruns);
    for (int q = 0; q < nruns; ++q) run_off[q] = pdep64((uint64_t)q << run_bits, tile_mask);

    auto* plan = new hq_group_plan();
    plan->L = L; plan->K = K; plan->NT = NT; plan->tile_mask = tile_mask;
    plan->meta = std::move(meta); plan->fixed_mask = fixed_mask; plan->fixed_value = fixed_value;
    plan->nrounds = nrounds; plan->nops = (int)dops.size(); plan->ngates = (int)hg.size(); plan->nlocal = nlocal;
    GroupParams& p = plan->p;
    p.ntiles = 1ull << (L - K - popcount64(fixed_mask));
    p.fixed_base = fixed_value;
    p.nruns = nruns;
    p.run_bytes = 16u << run_bits;
    p.nrounds = nrounds;
    p.nops = (int)dops.size();
    {   // tile number -> base: scatter over the runs of bits NOT in the tile
        const uint64_t outmask = ((1ull << L) - 1) & ~tile_mask & ~fixed_mask;
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
    plan->smem = (size_t)(16u << K) + 16 + (size_t)nrounds * sizeof(DevRound) + (dops.size() + 1) * sizeof(DevOp)   // +1: read-ahead slot
                 + (size_t)nrounds * NT * (8 + 4);                                                                 // gt, tb tables
    if (plan->smem > 227 * 1024) {
        delete plan;
        set_error("gate group does not fit in shared memory (too many gates/rounds for one launch): split it");
        return HQ_ERR_ARG;
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
        cudaError_t e = dev_alloc(&plan->dev_blob, total);
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
    // the specialised kernel's source (compiled at the first launch, or for a whole schedule at once by hq_group_plans_warm)
    if (rt().ready && jit_enabled()) plan->jit_identity = plan_identity(*plan);
    if (const char* dir = getenv("HQ_JIT_DUMP_DIR")) {   // developer aid: keep every emitted source (works without a GPU)
        static int serial = 0;
        const std::string src = jit_emit_source(*plan, false);
        char name[512];
        snprintf(name, sizeof(name), "%s/group_%03d.cu", dir, serial++);
        if (FILE* f = fopen(name, "w")) { fputs(src.c_str(), f); fclose(f); }
    }
    *out = plan;
    return HQ_OK;
}

// Everything the emitter reads from a plan, as bytes: two plans with the same identity get the same kernel, so the cache is
// keyed on this and a hit never emits any source.
static std::string plan_identity(const hq_group_plan& plan) {
    std::string id;
    auto put = [&](const void* p, size_t n) { id.append(static_cast<const char*>(p), n); };
    const int hdr[6] = {plan.L, plan.K, plan.NT, plan.nrounds, plan.nops, jit_l2_prefetch_slots()};
    put(hdr, sizeof(hdr));
    put(&plan.tile_mask, 8);
    put(&plan.fixed_mask, 8);
    put(&plan.fixed_value, 8);
    const char fuse = getenv("HQ_JIT_NO_FUSE") ? 0 : 1;
    put(&fuse, 1);
    for (const auto& m : plan.meta) {
        put(m.reg, sizeof(m.reg));
        for (int b : m.tbits) put(&b, sizeof(int));
    }
    put(plan.blob.data() + plan.o_rounds, (size_t)plan.nrounds * sizeof(DevRound));
    put(plan.blob.data() + plan.o_ops, (size_t)plan.nops * sizeof(DevOp));
    return id;
}

// Resolve the specialised kernel of a plan: memory cache, disk cache, or an NVRTC compile.
static void resolve_jit(const hq_group_plan* plan) {
    if (plan->jit || plan->jit_failed || plan->jit_identity.empty()) return;
    std::string why;
    const size_t smem = jit_smem_bytes(plan->K);
    JitKernel* k = jit_get(plan->jit_identity, smem, [&] { return jit_emit_source(*plan, false); }, &why);
    if (!k) {
        plan->jit_failed = true;
        static bool warned = false;
        if (!warned) { fprintf(stderr, "[hyquas_b200] specialised gate-group kernels unavailable (%s); using the interpreter kernel\n", why.c_str()); warned = true; }
        return;
    }
    plan->jit = k;
    plan->jit_occupancy = jit_max_blocks_per_sm(k, 2 * plan->NT, smem);
}

static void resolve_jit_zero(const hq_group_plan* plan) {
    if (plan->jit_zero || plan->jit_zero_failed || plan->jit_identity_zero.empty()) return;
    std::string why;
    JitKernel* k = jit_get(plan->jit_identity_zero, jit_smem_bytes(plan->K), [&] { return jit_emit_source(*plan, false, true); }, &why);
    if (!k) { plan->jit_zero_failed = true; return; }
    plan->jit_zero = k;
}

// The first gate group of a circuit starts from |0...0>: its zero-input variant writes every tile without reading any and makes
// the zero fill of the state unnecessary (one write sweep + one read sweep less per run).  Full-state plans only.
extern "C" int hq_group_plan_enable_zero_input(hq_group_plan* plan) {
    HQ_REQUIRE(plan != nullptr, "null plan");
    if (plan->jit_identity.empty() || plan->fixed_mask != 0) return HQ_OK;   // interpreter-only or per-chunk plan: not available
    plan->jit_identity_zero = plan->jit_identity + "|zero-input";
    return HQ_OK;
}

// Launch on a state whose contents are irrelevant: the result is the group applied to |0...0> (has_amp0: this rank holds amplitude 0).
// HQ_ERR_UNSUPPORTED when the variant is not available (the caller zero-fills and launches normally).
extern "C" int hq_group_plan_launch_from_zero(const hq_group_plan* plan, void* state, int has_amp0) {
    HQ_REQUIRE(plan != nullptr && state != nullptr, "null plan or state");
    HQ_REQUIRE(rt().ready, "hq_init() has not been called");
    resolve_jit_zero(plan);
    if (!plan->jit_zero) { set_error("zero-input variant not available for this plan"); return HQ_ERR_UNSUPPORTED; }
    const int occ = jit_max_blocks_per_sm(static_cast<JitKernel*>(plan->jit_zero), 2 * plan->NT, jit_smem_bytes(plan->K));
    plan->grid = (int)std::min<uint64_t>((plan->p.ntiles + 1) / 2, (uint64_t)std::max(1, (rt().sm_count - rt().reserved_ctas) * occ));
    return jit_launch(static_cast<JitKernel*>(plan->jit_zero), plan->grid, 2 * plan->NT, jit_smem_bytes(plan->K), rt().compute, state, has_amp0);
}

extern "C" int hq_group_plans_warm(hq_group_plan* const* plans, int n) {
    HQ_REQUIRE(n >= 0 && (n == 0 || plans != nullptr), "bad plan list");
    if (!rt().ready || !jit_enabled()) return HQ_OK;
    std::vector<std::string> ids;
    std::vector<const hq_group_plan*> which;
    for (int i = 0; i < n; ++i)
        if (plans[i] && !plans[i]->jit && !plans[i]->jit_failed && !plans[i]->jit_identity.empty() && !jit_cached(plans[i]->jit_identity)) {
            ids.push_back(plans[i]->jit_identity);
            which.push_back(plans[i]);
        }
    std::vector<char> zero(ids.size(), 0);
    for (int i = 0; i < n; ++i)
        if (plans[i] && !plans[i]->jit_zero && !plans[i]->jit_zero_failed && !plans[i]->jit_identity_zero.empty() && !jit_cached(plans[i]->jit_identity_zero)) {
            ids.push_back(plans[i]->jit_identity_zero);
            which.push_back(plans[i]);
            zero.push_back(1);
        }
    if (!ids.empty()) jit_precompile(ids.data(), (int)ids.size(), [&](int i) { return jit_emit_source(*which[i], false, zero[i] != 0); });
    for (int i = 0; i < n; ++i) if (plans[i]) { resolve_jit(plans[i]); resolve_jit_zero(plans[i]); }
    return HQ_OK;
}

extern "C" int hq_group_plan_is_specialised(const hq_group_plan* plan, int* yes) {
    HQ_REQUIRE(plan != nullptr && yes != nullptr, "null argument");
    *yes = plan->jit != nullptr;
    return HQ_OK;
}

template <int K, int MINB>
static int launch_k(const hq_group_plan* plan, GroupParams p, cudaStream_t s) {
    static bool attr_set = false;
    static std::map<size_t, int> occupancy;   // dynamic smem bytes -> resident CTAs per SM
    auto kern = group_kernel<K, MINB>;
    if (!attr_set) {
        HQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    auto it = occupancy.find(plan->smem);
    if (it == occupancy.end()) {
        int nb = 0;
        HQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, plan->NT, plan->smem));
        it = occupancy.emplace(plan->smem, std::max(1, nb)).first;
    }
    // an exchange kernel in flight owns `reserved_ctas` whole SMs: leave them out of the grid (no second wave)
    plan->grid = (int)std::min<uint64_t>(p.ntiles, (uint64_t)std::max(1, (rt().sm_count - rt().reserved_ctas) * it->second));
    kern<<<plan->grid, plan->NT, plan->smem, s>>>(p);
    HQ_CUDA(cudaGetLastError());
    return HQ_OK;
}

extern "C" int hq_group_plan_launch(const hq_group_plan* plan, void* state, int on_comm_stream) {
    HQ_REQUIRE(plan != nullptr && state != nullptr, "null plan or state");
    HQ_REQUIRE(rt().ready && plan->dev_blob != nullptr, "plan was created without a bound GPU (call hq_init first)");
    GroupParams p = plan->p;
    p.state = static_cast<double2*>(state);
    cudaStream_t s = on_comm_stream ? rt().comm : rt().compute;
    resolve_jit(plan);
    if (plan->jit) {
        // a CTA = two workers: at least two tiles per CTA when there are enough of them
        plan->grid = (int)std::min<uint64_t>((p.ntiles + 1) / 2, (uint64_t)std::max(1, (rt().sm_count - rt().reserved_ctas) * plan->jit_occupancy));
        return jit_launch(static_cast<JitKernel*>(plan->jit), plan->grid, 2 * plan->NT, jit_smem_bytes(plan->K), s, state);
    }
    const bool relaxed = rt().relaxed_regs;   // fewer resident CTAs, no register cap (HQ_RELAXED_REGS=1)
    // resident CTAs per SM asked of ptxas: 2048 threads' worth of registers at 64 (RBITS=3) / 128 (RBITS=4) per thread
    constexpr int B12 = RBITS == 4 ? 2 : 2, B11 = 2 * B12, B10 = 4 * B12;
    switch (plan->K) {
        case 10: return relaxed ? launch_k<10, B10 * 3 / 4>(plan, p, s) : launch_k<10, B10>(plan, p, s);
        case 11: return relaxed ? launch_k<11, B11 * 3 / 4>(plan, p, s) : launch_k<11, B11>(plan, p, s);
        case 12: return relaxed ? launch_k<12, 1>(plan, p, s) : launch_k<12, B12>(plan, p, s);
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

extern "C" int hq_group_plan_local_exchanges(const hq_group_plan* plan, int* n) {
    HQ_REQUIRE(plan != nullptr && n != nullptr, "null argument");
    *n = plan->nlocal;
    return HQ_OK;
}

extern "C" int hq_group_plan_table_bytes(const hq_group_plan* plan, int* bytes) {
    HQ_REQUIRE(plan != nullptr && bytes != nullptr, "null argument");
    *bytes = (int)plan->blob.size();
    return HQ_OK;
}

extern "C" int hq_group_plan_destroy(hq_group_plan* plan) {
    if (!plan) return HQ_OK;
    if (plan->dev_blob) dev_free(plan->dev_blob);
    if (plan->jit) jit_release(static_cast<JitKernel*>(plan->jit));
    if (plan->jit_zero) jit_release(static_cast<JitKernel*>(plan->jit_zero));
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
