// Fused dense-matrix (TransMM-class) kernel for sm_100a.
//
// What it computes is what the reference's BLAS groups compute (Executor::applyBlasGroup, src/executor.cpp:533-551):
// for a set of m qubits, every 2^m-vector of amplitudes obtained by fixing all other index bits is multiplied by a
// dense 2^m x 2^m complex matrix U (built on the host from the group's gates, GateGroup::initCPUMatrix,
// src/schedule.cpp:578-702).  The reference does it as cuTT transpose (state -> second buffer, so that the m qubits
// become the fastest index) followed by cublasZgemm (buffer -> state): two sweeps and twice the memory.  Here it is
// ONE in-place sweep:
//   * persistent CTAs loop over tiles of 2^Kt amplitudes (the m matrix bits + the low physical bits + padding);
//     a tile is gathered with 16-byte cp.async copies straight into an XOR-swizzled shared-memory image -- this is
//     the "qubit-permutation transpose", done on the fly by the address map instead of by a separate pass;
//   * the product runs on the FP64 tensor cores: mma.sync.aligned.m8n8k4.f64 (DMMA; tcgen05 has no f64 kind),
//     complex arithmetic as 4 real MMAs, U pre-arranged on the host in fragment order so that A-fragment loads are
//     linear, X fragments loaded conflict-free thanks to a per-plan swizzle chosen on the host;
//   * each warp owns whole columns of the tile (all 2^m rows of a few vectors), so results overwrite the inputs in
//     shared memory without any block-level barrier inside a matrix;
//   * several matrices on disjoint or overlapping qubit sets of the SAME tile can be chained in one launch
//     (barrier between them) -- the dense analogue of a gate group, amortising the HBM sweep;
//   * write-back with coalesced 128-bit stores.
// Algorithmic traffic: 32 bytes per amplitude per launch; flops: 8 * 2^m per amplitude per matrix.
#include <algorithm>
#include <complex>
#include <cstring>
#include <vector>

#include "hq_internal.h"

namespace hq {

constexpr int DENSE_MAX_MATS = 8;
constexpr int DENSE_THREADS = 256;

struct DenseMatDesc {
    int32_t m;            // matrix qubits (3..6)
    int32_t nblocks;      // column blocks of 8*NT vectors in one tile
    uint32_t u_off;       // first double2 of this matrix' fragment-ordered U in the U area
    uint32_t tab_off;     // first uint16 of {swk[2^m], swn[2^(Kt-m)]} in the table area
};

struct DenseParams {
    double2* state;
    uint64_t ntiles;
    uint64_t fixed_base;        // value of the fixed (non-varying) physical bits of this launch
    const double2* u_frag;      // all matrices, fragment order
    const uint16_t* tables;     // all matrices: swk | swn
    uint64_t g_high[16];        // physical offset contributed by bits 8..11 of the tile index
    uint16_t f_high[16];        // swizzle contribution of bits 8..11 of the tile index (already includes i << 8)
    uint64_t low_mask[8];       // physical bit (as a mask) of tile bits 0..7
    uint8_t fvec[12];           // 3-bit swizzle vector of each tile bit (bits 0..2: identity)
    int32_t Kt;
    int32_t nmat;
    uint32_t u_total;           // double2 count
    uint32_t tab_total;         // uint16 count
    int32_t nseg;
    uint8_t seg_shift[24], seg_src[24];
    uint64_t seg_mask[24];
    DenseMatDesc mats[DENSE_MAX_MATS];
};

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// One column block: RT row tiles (2^m = 8*RT rows) x NT column tiles (8 vectors each), all K = 8*RT inner steps.
template <int RT, int NT>
__device__ __forceinline__ void dense_block(double2* tile, const double2* __restrict__ ufrag, const uint16_t* __restrict__ swk,
                                            const uint16_t* __restrict__ swn, int n0, int lane) {
    constexpr int K4 = RT * 2;   // K / 4
    double accr[RT][NT][2], acci[RT][NT][2];
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
        for (int t = 0; t < NT; ++t) accr[r][t][0] = accr[r][t][1] = acci[r][t][0] = acci[r][t][1] = 0.0;
    uint32_t ncol[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) ncol[t] = swn[n0 + t * 8 + (lane >> 2)];
#pragma unroll 2
    for (int k4 = 0; k4 < K4; ++k4) {
        const uint32_t krow = swk[k4 * 4 + (lane & 3)];
        double2 x[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) x[t] = tile[krow ^ ncol[t]];
        // Two passes so that consecutive MMAs never hit the same accumulator back to back (asm volatile keeps this order):
        // first the Re(U) products of every (row tile, column tile), then the Im(U) products.
        double2 u[RT];
#pragma unroll
        for (int r = 0; r < RT; ++r) u[r] = ufrag[(r * K4 + k4) * 32 + lane];
#pragma unroll
        for (int r = 0; r < RT; ++r) {
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                dmma(accr[r][t][0], accr[r][t][1], u[r].x, x[t].x);
                dmma(acci[r][t][0], acci[r][t][1], u[r].x, x[t].y);
            }
        }
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            const double nui = -u[r].y;
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                dmma(accr[r][t][0], accr[r][t][1], nui, x[t].y);
                dmma(acci[r][t][0], acci[r][t][1], u[r].y, x[t].x);
            }
        }
    }
    // every lane's X fragments were consumed by the (warp-synchronous) MMAs above: the columns can be overwritten
#pragma unroll
    for (int r = 0; r < RT; ++r) {
        const uint32_t krow = swk[r * 8 + (lane >> 2)];
#pragma unroll
        for (int t = 0; t < NT; ++t) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const uint32_t col = swn[n0 + t * 8 + (lane & 3) * 2 + c];
                tile[krow ^ col] = make_double2(accr[r][t][c], acci[r][t][c]);
            }
        }
    }
}

// MAXM = 4: every matrix of the launch has <= 4 qubits (16 accumulator doubles per thread, 3 CTAs per SM);
// MAXM = 6: general (32 accumulator doubles, 2 CTAs per SM).
// DB: two tile buffers per CTA (one CTA per SM); the gather of tile i+1 is issued BEFORE tile i is computed, so HBM
// reads run under the tensor-core phase of the same CTA by construction.  (With one buffer and 2-3 CTAs per SM the CTAs fall
// into lock step -- all loading, then all computing -- and the DMMA pipe idles a third of the time: ncu, profiles/r01_s10.)
template <int MAXM, bool DB>
__global__ void __launch_bounds__(DENSE_THREADS, DB ? 1 : (MAXM <= 4 ? 3 : 2)) dense_kernel(const __grid_constant__ DenseParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int TILE = 1 << P.Kt;
    double2* tile0 = reinterpret_cast<double2*>(smem_raw);
    double2* u_s = tile0 + (DB ? 2 : 1) * TILE;
    uint16_t* tab_s = reinterpret_cast<uint16_t*>(u_s + P.u_total);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (uint32_t i = tid; i < P.u_total; i += DENSE_THREADS) u_s[i] = __ldg(P.u_frag + i);
    for (uint32_t i = tid; i < P.tab_total; i += DENSE_THREADS) tab_s[i] = __ldg(P.tables + i);

    // this thread's share of a tile: elements j = tid + 256 * i.  Physical offset and swizzled position both split into
    // a per-thread part (bits 0..7 of j) and a per-i part (bits 8..11).
    uint64_t g_low = 0;
    uint32_t f_low = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b)
        if (tid >> b & 1) {
            g_low |= P.low_mask[b];
            if (b >= 3) f_low ^= P.fvec[b];
        }
    const uint32_t s_low = (uint32_t)tid ^ f_low;
    const int per_thread = TILE / DENSE_THREADS;
    const uint32_t tile0_s = (uint32_t)__cvta_generic_to_shared(tile0);
    __syncthreads();

    auto tile_base = [&](uint64_t t) {
        uint64_t base = P.fixed_base;
        for (int s = 0; s < P.nseg; ++s) base |= ((t >> P.seg_src[s]) & P.seg_mask[s]) << P.seg_shift[s];
        return P.state + base + g_low;
    };
    auto gather = [&](double2* gbase, int buf) {
        const uint32_t dst = tile0_s + (uint32_t)buf * (uint32_t)TILE * 16u;
        for (int i = 0; i < per_thread; ++i) cp_async16(dst + ((s_low ^ P.f_high[i]) << 4), gbase + P.g_high[i]);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    uint64_t t = blockIdx.x;
    int cur = 0;
    double2* gbase = nullptr;
    if (t < P.ntiles) {
        gbase = tile_base(t);
        gather(gbase, 0);
    }
    for (; t < P.ntiles; t += gridDim.x) {
        const uint64_t tn = t + gridDim.x;
        double2* gnext = nullptr;
        if (DB && tn < P.ntiles) {   // prefetch the next tile into the other buffer, then wait for the current one only
            gnext = tile_base(tn);
            gather(gnext, cur ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            cp_async_wait_all();
        }
        __syncthreads();
        double2* tile = tile0 + (size_t)cur * TILE;

        for (int mi = 0; mi < P.nmat; ++mi) {
            const DenseMatDesc d = P.mats[mi];
            const double2* uf = u_s + d.u_off;
            const uint16_t* swk = tab_s + d.tab_off;
            const uint16_t* swn = swk + (1 << d.m);
            for (int nb = warp; nb < d.nblocks; nb += DENSE_THREADS / 32) {
                if (MAXM <= 4) {
                    if (d.m == 3) dense_block<1, 4>(tile, uf, swk, swn, nb * 32, lane);
                    else dense_block<2, 2>(tile, uf, swk, swn, nb * 16, lane);
                } else {
                    switch (d.m) {
                        case 3: dense_block<1, 4>(tile, uf, swk, swn, nb * 32, lane); break;
                        case 4: dense_block<2, 2>(tile, uf, swk, swn, nb * 16, lane); break;
                        case 5: dense_block<4, 2>(tile, uf, swk, swn, nb * 16, lane); break;
                        default: dense_block<8, 1>(tile, uf, swk, swn, nb * 8, lane); break;
                    }
                }
            }
            __syncthreads();
        }

        for (int i = 0; i < per_thread; ++i) gbase[P.g_high[i]] = tile[s_low ^ P.f_high[i]];
        if (DB) {
            gbase = gnext;
            cur ^= 1;
            // the buffer just stored is refilled by the NEXT iteration's prefetch: every thread must be done reading it
            __syncthreads();
        } else {
            __syncthreads();
            if (tn < P.ntiles) {
                gbase = tile_base(tn);
                gather(gbase, 0);
            }
        }
    }
}

}  // namespace hq

using namespace hq;

struct hq_dense_plan {
    int L = 0, Kt = 0, nmat = 0;
    size_t smem = 0;
    mutable int grid = 0;
    double flops_per_amp = 0;
    std::vector<unsigned char> blob;   // u_frag | tables
    size_t o_tab = 0;
    void* dev_blob = nullptr;
    DenseParams p{};
};


// qubit_pos[i]: physical local bit of matrix qubit i (bit i of the row/column index of U);  U column-major
// (U[row + col * 2^m], like the A operand of the reference's cublasZgemm call), interleaved re/im.
extern "C" int hq_dense_plan_create(int L, int nmat, const int* m_list, const int* qubit_pos, const double* u_colmajor,
                                    hq_dense_plan** out) {
    return hq_dense_plan_create_ex(L, 0, 0, nmat, m_list, qubit_pos, u_colmajor, out);
}

extern "C" int hq_dense_plan_create_ex(int L, uint64_t fixed_mask, uint64_t fixed_value, int nmat, const int* m_list,
                                       const int* qubit_pos, const double* u_colmajor, hq_dense_plan** out) {
    HQ_REQUIRE((fixed_mask >> L) == 0 && (fixed_value & ~fixed_mask) == 0, "fixed bits must be local and fixed_value within fixed_mask");
    HQ_REQUIRE(out && nmat >= 1 && nmat <= DENSE_MAX_MATS && m_list && qubit_pos && u_colmajor, "bad arguments to hq_dense_plan_create");
    HQ_REQUIRE(L - popcount64(fixed_mask) >= 8 && L <= 40, "local qubit count out of range for the dense kernel");
    const int Kt = std::min(12, L - popcount64(fixed_mask));
    // ---- tile = union of all matrix bits + lowest free physical bits ----
    uint64_t tile_mask = 0;
    {
        const int* qp = qubit_pos;
        for (int i = 0; i < nmat; ++i) {
            HQ_REQUIRE(m_list[i] >= 1 && m_list[i] <= 6, "matrix size must be 1..6 qubits");
            uint64_t seen = 0;
            for (int b = 0; b < m_list[i]; ++b, ++qp) {
                HQ_REQUIRE(*qp >= 0 && *qp < L, "matrix qubit outside the local state");
                HQ_REQUIRE(!(fixed_mask >> *qp & 1), "matrix qubit on a fixed bit");
                HQ_REQUIRE(!(seen >> *qp & 1), "matrix qubits must be distinct");
                seen |= 1ull << *qp;
            }
            tile_mask |= seen;
        }
    }
    HQ_REQUIRE(popcount64(tile_mask) <= Kt - 3 || popcount64(tile_mask | 7ull) <= Kt, "matrices span too many qubits for one tile");
    tile_mask |= 7ull;   // >= 128-byte runs
    HQ_REQUIRE((fixed_mask & 7ull) == 0, "physical bits 0..2 cannot be fixed");
    for (int b = 0; b < L && popcount64(tile_mask) < Kt; ++b) if (!(fixed_mask >> b & 1)) tile_mask |= 1ull << b;
    HQ_REQUIRE(popcount64(tile_mask) == Kt, "matrices span too many qubits for one tile");
    int phys_to_tile[64];
    for (int i = 0; i < 64; ++i) phys_to_tile[i] = -1;
    int tile_to_phys[12];
    for (int b = 0, k = 0; b < L; ++b)
        if (tile_mask >> b & 1) { phys_to_tile[b] = k; tile_to_phys[k] = b; ++k; }

    // ---- per-matrix index orders + one swizzle for the whole tile ----
    // Swizzle: position of tile index j = j ^ f(j), f(j) = XOR of fvec[b] over set bits b >= 3 (bits 0..2 contribute
    // themselves).  A quarter-warp of a 128-bit shared access is conflict-free iff the bank vectors (fvec, or the unit
    // vectors for b < 3) of the three tile bits that vary inside it are linearly independent.  The X-fragment load varies
    // (k bit 0, k bit 1, n bit 0) and the result store varies (k bit 0, n bit 1, n bit 2): choose index orders and fvec so
    // that both triples are independent for every matrix where possible.
    uint8_t fvec[12];
    bool fixed[12];
    for (int b = 0; b < 12; ++b) { fvec[b] = b < 3 ? (uint8_t)(1u << b) : 0; fixed[b] = b < 3; }
    struct MatPlan { int m, pad; std::vector<int> kbits, nbits; };   // tile bits in index order (bit 0 first)
    std::vector<MatPlan> mp(nmat);
    auto independent = [](uint8_t a, uint8_t b, uint8_t c) { return a && b && c && a != b && a != c && b != c && (a ^ b) != c; };
    {
        const int* qp = qubit_pos;
        for (int i = 0; i < nmat; ++i) {
            MatPlan& M = mp[i];
            M.m = m_list[i];
            std::vector<int> kb;
            for (int b = 0; b < M.m; ++b) kb.push_back(phys_to_tile[*qp++]);
            M.pad = 0;
            // matrices smaller than 8x8 are padded with identity on the lowest free tile bits (DMMA needs 8 rows)
            for (int b = 0; b < Kt && (int)kb.size() < 3; ++b)
                if (std::find(kb.begin(), kb.end(), b) == kb.end()) { kb.push_back(b); M.pad++; }
            std::vector<int> nb;
            for (int b = 0; b < Kt; ++b) if (std::find(kb.begin(), kb.end(), b) == kb.end()) nb.push_back(b);
            HQ_REQUIRE((int)nb.size() >= 5 || (int)kb.size() == 6, "tile too small for this matrix");
            // assign bank vectors greedily: the six "phase" bits of this matrix get vectors making both triples independent
            auto pick = [&](int bit, uint8_t avoid1, uint8_t avoid2) {   // choose fvec[bit] not in span{avoid1, avoid2}
                if (fixed[bit]) return;
                for (uint8_t v = 1; v < 8; ++v) {
                    if (v == avoid1 || v == avoid2 || v == (avoid1 ^ avoid2)) continue;
                    fvec[bit] = v;
                    break;
                }
                fixed[bit] = true;
            };
            // order: prefer already-fixed low bits first so that their unit vectors anchor the triples
            std::stable_sort(kb.begin(), kb.begin() + M.m, [&](int a, int b) { return fixed[a] > fixed[b]; });
            std::stable_sort(nb.begin(), nb.end(), [&](int a, int b) { return fixed[a] > fixed[b]; });
            // try a few rotations of the index orders to find one where the fixed vectors do not already clash
            bool done = false;
            for (int rk = 0; rk < (int)kb.size() && !done; ++rk) {
                for (int rn = 0; rn < (int)nb.size() && !done; ++rn) {
                    std::vector<int> k2 = kb, n2 = nb;
                    std::rotate(k2.begin(), k2.begin() + rk, k2.end());
                    std::rotate(n2.begin(), n2.begin() + rn, n2.end());
                    uint8_t save[12]; bool savef[12];
                    std::memcpy(save, fvec, 12); std::memcpy(savef, fixed, 12);
                    pick(k2[0], 0, 0);
                    pick(k2[1], fvec[k2[0]], 0);
                    pick(n2[0], fvec[k2[0]], fvec[k2[1]]);
                    pick(n2[1], fvec[k2[0]], 0);
                    pick(n2[2], fvec[k2[0]], fvec[n2[1]]);
                    if (independent(fvec[k2[0]], fvec[k2[1]], fvec[n2[0]]) && independent(fvec[k2[0]], fvec[n2[1]], fvec[n2[2]])) {
                        kb = k2; nb = n2; done = true;
                    } else {
                        std::memcpy(fvec, save, 12); std::memcpy(fixed, savef, 12);
                    }
                }
            }
            if (!done) {   // accept bank conflicts for this matrix (rare: several matrices with clashing needs)
                pick(kb[0], 0, 0); pick(kb[1], fvec[kb[0]], 0); pick(nb[0], fvec[kb[0]], fvec[kb[1]]);
                pick(nb[1], fvec[kb[0]], 0); pick(nb[2], fvec[kb[0]], fvec[nb[1]]);
            }
            M.kbits = kb;
            M.nbits = nb;
        }
        for (int b = 3; b < 12; ++b) if (!fixed[b]) fvec[b] = (uint8_t)(1u << (b % 3));
    }
    auto swz = [&](uint32_t j) {
        uint32_t f = 0;
        for (int b = 3; b < Kt; ++b) if (j >> b & 1) f ^= fvec[b];
        return j ^ f;
    };

    // ---- tables + fragment-ordered U ----
    std::vector<double2> ufrag;
    std::vector<uint16_t> tables;
    auto* plan = new hq_dense_plan();
    DenseParams& p = plan->p;
    std::memset(&p, 0, sizeof(p));
    const double* usrc = u_colmajor;
    const int* qp = qubit_pos;
    for (int i = 0; i < nmat; ++i) {
        const MatPlan& M = mp[i];
        const int m = M.m, me = (int)M.kbits.size(), K = 1 << me, Korig = 1 << m;
        // kbits (tile bits, in the index order chosen above) -> which ORIGINAL matrix qubit (or padding) each one is
        std::vector<int> orig_of(me, -1);
        for (int b = 0; b < me; ++b)
            for (int q = 0; q < m; ++q)
                if (phys_to_tile[qp[q]] == M.kbits[b]) orig_of[b] = q;
        auto to_orig = [&](int kk, int& pad_bits) {   // index in the new order -> index of the user's U, padding bits apart
            int o = 0; pad_bits = 0;
            for (int b = 0; b < me; ++b)
                if (kk >> b & 1) { if (orig_of[b] >= 0) o |= 1 << orig_of[b]; else pad_bits |= 1 << b; }
            return o;
        };
        DenseMatDesc& d = p.mats[i];
        d.m = me;
        d.u_off = (uint32_t)ufrag.size();
        d.tab_off = (uint32_t)tables.size();
        const int RT = K / 8, K4 = K / 4;
        const int NT = me <= 3 ? 4 : (me <= 5 ? 2 : 1);
        d.nblocks = (1 << (Kt - me)) / (8 * NT);
        HQ_REQUIRE(d.nblocks >= 1, "tile too small for this matrix");
        for (int r = 0; r < RT; ++r)
            for (int k4 = 0; k4 < K4; ++k4)
                for (int lane = 0; lane < 32; ++lane) {
                    const int row = r * 8 + lane / 4, col = k4 * 4 + lane % 4;
                    int prow, pcol;
                    const int orow = to_orig(row, prow), ocol = to_orig(col, pcol);
                    double2 v = make_double2(0.0, 0.0);
                    if (prow == pcol) v = make_double2(usrc[2 * ((size_t)orow + (size_t)ocol * Korig)], usrc[2 * ((size_t)orow + (size_t)ocol * Korig) + 1]);
                    ufrag.push_back(v);
                }
        for (int kk = 0; kk < K; ++kk) {
            uint32_t j = 0;
            for (int b = 0; b < me; ++b) if (kk >> b & 1) j |= 1u << M.kbits[b];
            tables.push_back((uint16_t)swz(j));
        }
        for (int nn = 0; nn < (1 << (Kt - me)); ++nn) {
            uint32_t j = 0;
            for (int b = 0; b < Kt - me; ++b) if (nn >> b & 1) j |= 1u << M.nbits[b];
            tables.push_back((uint16_t)swz(j));
        }
        plan->flops_per_amp += 8.0 * K;
        usrc += 2 * (size_t)Korig * Korig;
        qp += m;
    }
    plan->L = L; plan->Kt = Kt; plan->nmat = nmat;
    p.Kt = Kt; p.nmat = nmat;
    p.ntiles = 1ull << (L - Kt - popcount64(fixed_mask));
    p.fixed_base = fixed_value;
    p.u_total = (uint32_t)ufrag.size();
    p.tab_total = (uint32_t)tables.size();
    std::memcpy(p.fvec, fvec, 12);
    for (int b = 0; b < 8; ++b) p.low_mask[b] = b < Kt ? 1ull << tile_to_phys[b] : 0;
    for (int i = 0; i < 16; ++i) {
        uint64_t g = 0; uint32_t j = (uint32_t)i << 8;
        for (int b = 8; b < Kt; ++b) if (j >> b & 1) g |= 1ull << tile_to_phys[b];
        p.g_high[i] = g;
        p.f_high[i] = (uint16_t)(swz(j));
    }
    {
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
    plan->smem = ((size_t)16 << Kt) + ufrag.size() * 16 + ((tables.size() * 2 + 15) & ~size_t(15));
    if (plan->smem > 227 * 1024) {
        delete plan;
        set_error("dense group does not fit in shared memory: fewer / smaller matrices per launch");
        return HQ_ERR_ARG;
    }
    plan->o_tab = ufrag.size() * 16;
    plan->blob.assign(plan->o_tab + ((tables.size() * 2 + 15) & ~size_t(15)), 0);
    std::memcpy(plan->blob.data(), ufrag.data(), ufrag.size() * 16);
    std::memcpy(plan->blob.data() + plan->o_tab, tables.data(), tables.size() * 2);
    if (rt().ready) {
        cudaError_t e = dev_alloc(&plan->dev_blob, plan->blob.size());
        if (e == cudaSuccess) e = cudaMemcpyAsync(plan->dev_blob, plan->blob.data(), plan->blob.size(), cudaMemcpyHostToDevice, rt().compute);
        if (e == cudaSuccess) e = cudaStreamSynchronize(rt().compute);
        if (e != cudaSuccess) { delete plan; return cuda_fail(e, "dense plan upload", __FILE__, __LINE__); }
        p.u_frag = reinterpret_cast<const double2*>(plan->dev_blob);
        p.tables = reinterpret_cast<const uint16_t*>(static_cast<unsigned char*>(plan->dev_blob) + plan->o_tab);
    }
    *out = plan;
    return HQ_OK;
}

extern "C" int hq_dense_plan_launch(const hq_dense_plan* plan, void* state, int on_comm_stream) {
    HQ_REQUIRE(plan && state, "null plan or state");
    HQ_REQUIRE(rt().ready && plan->dev_blob, "dense plan was created without a bound GPU (call hq_init first)");
    static bool attr_set = false;
    static int want_db = -1;   // HQ_DENSE_DB: unset = per plan (below), 0 = never, 1 = whenever it fits
    if (!attr_set) {
        HQ_CUDA(cudaFuncSetAttribute(dense_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        HQ_CUDA(cudaFuncSetAttribute(dense_kernel<6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        HQ_CUDA(cudaFuncSetAttribute(dense_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        HQ_CUDA(cudaFuncSetAttribute(dense_kernel<6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        if (const char* e = getenv("HQ_DENSE_DB")) want_db = atoi(e) != 0 ? 1 : 0;
        attr_set = true;
    }
    int maxm = 0;
    for (int i = 0; i < plan->p.nmat; ++i) maxm = std::max(maxm, plan->p.mats[i].m);
    const size_t smem_db = plan->smem + ((size_t)16 << plan->Kt);
    // Which shape?  Measured at 2^30 amplitudes (profiles/r01_s10_microbench.json = 3 CTAs/SM single buffer, r01_s11 = double
    // buffer): the double buffer wins where one matrix dominates the tile's time (m = 6: 18.8 vs 20.9 ms; a lone m <= 3:
    // 5.4 vs 5.9 ms) and loses where several small matrices share a launch (m4: 7.5 vs 6.0, m4x2: 12.1 vs 10.4, m3x3: 10.6 vs
    // 8.8) -- eight warps cannot cover the barriers between matrices.
    const bool db_pays = maxm >= 6 || (plan->p.nmat == 1 && maxm <= 3);
    const bool db = (want_db < 0 ? db_pays : want_db != 0) && smem_db <= 227 * 1024;
    auto kern = db ? (maxm <= 4 ? dense_kernel<4, true> : dense_kernel<6, true>) : (maxm <= 4 ? dense_kernel<4, false> : dense_kernel<6, false>);
    const size_t smem = db ? smem_db : plan->smem;
    int nb = 0;
    HQ_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, DENSE_THREADS, smem));
    DenseParams p = plan->p;
    p.state = static_cast<double2*>(state);
    // an exchange kernel in flight owns rt().reserved_ctas whole SMs (fat CTAs, see swap.cu): leave them out of the grid
    const int reserve = rt().reserved_ctas * std::max(1, nb);
    plan->grid = (int)std::min<uint64_t>(p.ntiles, (uint64_t)std::max(1, rt().sm_count * std::max(1, nb) - reserve));
    kern<<<plan->grid, DENSE_THREADS, smem, on_comm_stream ? rt().comm : rt().compute>>>(p);
    HQ_CUDA(cudaGetLastError());
    return HQ_OK;
}

extern "C" int hq_dense_plan_info(const hq_dense_plan* plan, int* tile_bits, int* smem_bytes, int* grid, double* flops_per_amp, int* table_bytes) {
    HQ_REQUIRE(plan != nullptr, "null plan");
    if (tile_bits) *tile_bits = plan->Kt;
    if (smem_bytes) *smem_bytes = (int)plan->smem;
    if (grid) *grid = plan->grid;
    if (flops_per_amp) *flops_per_amp = plan->flops_per_amp;
    if (table_bytes) *table_bytes = (int)plan->blob.size();
    return HQ_OK;
}

extern "C" int hq_dense_plan_destroy(hq_dense_plan* plan) {
    if (!plan) return HQ_OK;
    if (plan->dev_blob) dev_free(plan->dev_blob);
    delete plan;
    return HQ_OK;
}

extern "C" int hq_dense_apply(void* state, int L, int m, const int* qubit_pos, const double* u_colmajor) {
    hq_dense_plan* plan = nullptr;
    int rc = hq_dense_plan_create(L, 1, &m, qubit_pos, u_colmajor, &plan);
    if (rc != HQ_OK) return rc;
    rc = hq_dense_plan_launch(plan, state, 0);
    if (rc == HQ_OK) {
        cudaError_t e = cudaStreamSynchronize(rt().compute);
        if (e != cudaSuccess) rc = cuda_fail(e, "dense sync", __FILE__, __LINE__);
    }
    hq_dense_plan_destroy(plan);
    return rc;
}

// TEST HOOK (CPU suite): interpret the plan's device tables on a host array, mirroring dense_kernel serially.
extern "C" int hq_debug_dense_plan_emulate(const hq_dense_plan* plan, double* state_re_im) {
    HQ_REQUIRE(plan && state_re_im, "null argument");
    const DenseParams& p = plan->p;
    const int Kt = p.Kt, TILE = 1 << Kt;
    const double2* ufrag = reinterpret_cast<const double2*>(plan->blob.data());
    const uint16_t* tables = reinterpret_cast<const uint16_t*>(plan->blob.data() + plan->o_tab);
    double2* st = reinterpret_cast<double2*>(state_re_im);
    std::vector<double2> tile(TILE), y;
    for (uint64_t t = 0; t < p.ntiles; ++t) {
        uint64_t base = p.fixed_base;
        for (int s = 0; s < p.nseg; ++s) base |= ((t >> p.seg_src[s]) & p.seg_mask[s]) << p.seg_shift[s];
        std::vector<uint64_t> goff(TILE);
        std::vector<uint32_t> spos(TILE);
        for (int tid = 0; tid < DENSE_THREADS; ++tid) {
            uint64_t g_low = 0; uint32_t f_low = 0;
            for (int b = 0; b < 8; ++b) if (tid >> b & 1) { g_low |= p.low_mask[b]; if (b >= 3) f_low ^= p.fvec[b]; }
            for (int i = 0; i < TILE / DENSE_THREADS; ++i) {
                const int j = tid + DENSE_THREADS * i;
                goff[j] = base + g_low + p.g_high[i];
                spos[j] = ((uint32_t)tid ^ f_low) ^ p.f_high[i];
                tile[spos[j]] = st[goff[j]];
            }
        }
        for (int mi = 0; mi < p.nmat; ++mi) {
            const DenseMatDesc& d = p.mats[mi];
            const int K = 1 << d.m, K4 = K / 4, N = 1 << (Kt - d.m);
            const uint16_t* swk = tables + d.tab_off;
            const uint16_t* swn = swk + K;
            const double2* uf = ufrag + d.u_off;
            y.assign(K, make_double2(0, 0));
            for (int n = 0; n < N; ++n) {
                for (int row = 0; row < K; ++row) {
                    double ar = 0, ai = 0;
                    for (int col = 0; col < K; ++col) {
                        const double2 u = uf[((row / 8) * K4 + col / 4) * 32 + (row % 8) * 4 + col % 4];
                        const double2 x = tile[swk[col] ^ swn[n]];
                        ar += u.x * x.x - u.y * x.y;
                        ai += u.x * x.y + u.y * x.x;
                    }
                    y[row] = make_double2(ar, ai);
                }
                for (int row = 0; row < K; ++row) tile[swk[row] ^ swn[n]] = y[row];
            }
        }
        for (int j = 0; j < TILE; ++j) st[goff[j]] = tile[spos[j]];
    }
    return HQ_OK;
}
