// TEST HOOK -- never called by the product path (Circuit::run / Executor only ever launch the CUDA kernel).
//
// hq_debug_group_plan_emulate() walks the *device tables* of a gate-group plan (run offsets, tile-base
// segments, per-round thread/register index tables, the lowered op list) on a host array, mirroring
// group_kernel<K> instruction for instruction but serially.  It exists so that the `-m "not gpu"` tests can
// prove the planner's encoding against the oracle on a machine without a GPU; GPU parity tests go through
// hq_group_plan_launch.  It is deliberately slow and is not exported in include/hyquas_b200.h's product section.
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "group_plan.h"
#include "hq_internal.h"

using namespace hq;

namespace {
struct Amp { double x, y; };

void cmul(Amp& v, double fr, double fi) { Amp x = v; v.x = fr * x.x - fi * x.y; v.y = fr * x.y + fi * x.x; }

// returns how many extra entries were consumed (diagonal runs)
int applyOp(Amp* a, const DevOp* op, uint64_t phys) {
    const DevOp& o = *op;
    if (o.code == CODE_DIAG_RUN) {
        double fr = 1.0, fi = 0.0;
        for (uint32_t e = 1; e <= o.aux; ++e) {
            const DevOp& d = op[e];
            if ((phys & d.cphys) != d.cphys) continue;
            const bool hi = (d.tphys == 0) || (phys & d.tphys);
            if (!hi && (d.flags & 1u)) continue;
            const double dr = hi ? d.m[6] : d.m[0], di = hi ? d.m[7] : d.m[1];
            const double nr = fr * dr - fi * di;
            fi = fi * dr + fr * di;
            fr = nr;
        }
        for (int i = 0; i < R; ++i)
            if ((i & o.creg) == o.creg) cmul(a[i], fr, fi);
        return (int)o.aux;
    }
    if ((phys & o.cphys) != o.cphys) return 0;
    const uint32_t creg = o.creg;
    if (o.code == CODE_DIAG_T) {
        const bool hi = (o.tphys == 0) || (phys & o.tphys);
        if (!hi && (o.flags & 1u)) return 0;
        for (int i = 0; i < R; ++i)
            if ((i & creg) == creg) cmul(a[i], hi ? o.m[6] : o.m[0], hi ? o.m[7] : o.m[1]);
        return 0;
    }
    const uint32_t kind = o.code / 24, tb = (o.code % 24) / 6, cbc = o.code % 6;
    // the templated control case must agree with the full mask the planner stored
    const uint32_t expect = cbc == 0 ? 0u : (cbc <= 4 ? 1u << (cbc - 1) : creg);
    if (expect != creg) { std::fprintf(stderr, "plan emulator: control case %u disagrees with mask %u\n", cbc, creg); std::abort(); }
    for (int p = 0; p < R / 2; ++p) {
        const int lo = ((p >> tb) << (tb + 1)) | (p & ((1 << tb) - 1)), hi = lo | (1 << tb);
        if ((lo & creg) != creg) continue;
        const Amp x = a[lo], y = a[hi];
        const double* m = o.m;
        switch (kind) {
            case OP_GEN: {   // m = the complex matrix [[a, b], [c, d]]
                const std::complex<double> A(m[0], m[1]), B(m[2], m[3]), C(m[4], m[5]), D(m[6], m[7]);
                const std::complex<double> lo_(x.x, x.y), hi_(y.x, y.y);
                const std::complex<double> nl = A * lo_ + B * hi_, nh = C * lo_ + D * hi_;
                a[lo] = {nl.real(), nl.imag()}; a[hi] = {nh.real(), nh.imag()};
                break;
            }
            case OP_REAL: {  // m = {a, b, c, d}: real [[a, b], [c, d]]
                a[lo] = {m[0] * x.x + m[1] * y.x, m[0] * x.y + m[1] * y.y};
                a[hi] = {m[2] * x.x + m[3] * y.x, m[2] * x.y + m[3] * y.y};
                break;
            }
            case OP_RXL: {   // m = {a, b, c, d} of [[a, i b], [i c, d]]
                a[lo] = {m[0] * x.x - m[1] * y.y, m[0] * x.y + m[1] * y.x};
                a[hi] = {-m[2] * x.y + m[3] * y.x, m[2] * x.x + m[3] * y.y};
                break;
            }
            case OP_SWAP: a[lo] = y; a[hi] = x; break;
            case OP_YL: a[lo] = {y.y, -y.x}; a[hi] = {-x.y, x.x}; break;
            case OP_DIAG_R:
                cmul(a[lo], m[0], m[1]);
                cmul(a[hi], m[6], m[7]);
                break;
            case OP_DIAG_R1: cmul(a[hi], m[6], m[7]); break;
            case OP_ZFLIP: a[hi] = {-y.x, -y.y}; break;
            default: {
                if (kind >= OP_BF0 && kind <= OP_BF7) {   // [[1, p],[q, -p q]] by definition (the PTX uses add/sub butterflies)
                    const double* pq = HQ_BF_PQ[kind - OP_BF0];
                    const std::complex<double> pp(pq[0], pq[1]), qq(pq[2], pq[3]), lo_(x.x, x.y), hi_(y.x, y.y);
                    const std::complex<double> nl = lo_ + pp * hi_, nh = qq * lo_ - pp * qq * hi_;
                    a[lo] = {nl.real(), nl.imag()}; a[hi] = {nh.real(), nh.imag()};
                    break;
                }
                std::fprintf(stderr, "plan emulator: bad op code %u\n", o.code); std::abort();
            }
        }
    }
    return 0;
}
}  // namespace

extern "C" int hq_debug_group_plan_emulate(const hq_group_plan* plan, double* state_re_im) {
    HQ_REQUIRE(plan && state_re_im, "null plan or state");
    Amp* state = reinterpret_cast<Amp*>(state_re_im);
    const unsigned char* b = plan->blob.data();
    const uint64_t* run_off = reinterpret_cast<const uint64_t*>(b + plan->o_run);
    const DevRound* rounds = reinterpret_cast<const DevRound*>(b + plan->o_rounds);
    const DevOp* ops = reinterpret_cast<const DevOp*>(b + plan->o_ops);
    const uint64_t* gt = reinterpret_cast<const uint64_t*>(b + plan->o_gt);
    const uint16_t* tb = reinterpret_cast<const uint16_t*>(b + plan->o_tb);
    const GroupParams& P = plan->p;
    const int NT = plan->NT, TILE = 1 << plan->K;
    const uint32_t run_amps = P.run_bytes >> 4;
    std::vector<Amp> sm(TILE), regs((size_t)NT * R);
    for (uint64_t t = 0; t < P.ntiles; ++t) {
        uint64_t base = P.fixed_base;
        for (int s = 0; s < P.nseg; ++s) base |= ((t >> P.seg_src[s]) & P.seg_mask[s]) << P.seg_shift[s];
        for (int q = 0; q < P.nruns; ++q)   // the producer warp's bulk copies
            std::memcpy(&sm[(size_t)q * run_amps], &state[base + run_off[q]], P.run_bytes);
        for (int r = 0; r < P.nrounds; ++r) {
            const DevRound& rd = rounds[r];
            for (int tid = 0; tid < NT; ++tid) {   // read phase of every consumer thread
                const uint32_t tin = tb[(size_t)(2 * r) * NT + tid];
                for (int i = 0; i < R; ++i) regs[(size_t)tid * R + i] = sm[tin ^ rd.ro_in[i]];
            }
            for (int tid = 0; tid < NT; ++tid) {
                Amp* a = &regs[(size_t)tid * R];
                const uint64_t phys = base | gt[(size_t)r * NT + tid];
                for (int op = rd.op_begin; op < rd.op_end; ++op) op += applyOp(a, &ops[op], phys);
                if (rd.flags & 2u) {
                    for (int i = 0; i < R; ++i) state[phys + rd.go[i]] = a[i];
                } else {
                    const uint32_t tout = tb[(size_t)(2 * r + 1) * NT + tid];
                    for (int i = 0; i < R; ++i) sm[tout ^ rd.ro_out[i]] = a[i];
                }
            }
        }
    }
    return HQ_OK;
}
