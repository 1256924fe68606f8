// TEST HOOK -- never called by the product path (Circuit::run / Executor only ever launch the CUDA kernel).
//
// hq_debug_group_plan_emulate() walks the *device tables* of a gate-group plan (run offsets, tile-base
// segments, per-round thread/register index tables, the lowered op list) on a host array, mirroring
// group_kernel<K> instruction for instruction but serially.  It exists so that the `-m "not gpu"` tests can
// prove the planner's encoding against the oracle on a machine without a GPU; GPU parity tests go through
// hq_group_plan_launch.  It is deliberately slow and is not exported in include/hyquas_b200.h's product section.
#include <cstring>
#include <vector>

#include "group_plan.h"
#include "hq_internal.h"

using namespace hq;

namespace {
struct Amp { double x, y; };

void applyOp(Amp* a, const DevOp& o, uint64_t phys) {
    if ((phys & o.cphys) != o.cphys) return;
    const uint32_t creg = o.creg;
    if (o.kind == OP_DIAG_T) {
        const bool hi = (o.tphys == 0) || (phys & o.tphys);
        if (!hi && (o.flags & 1u)) return;
        const double fr = hi ? o.m[6] : o.m[0], fi = hi ? o.m[7] : o.m[1];
        for (int i = 0; i < R; ++i)
            if ((i & creg) == creg) { Amp x = a[i]; a[i].x = fr * x.x - fi * x.y; a[i].y = fr * x.y + fi * x.x; }
        return;
    }
    const int tb = (int)o.tbit;
    for (int p = 0; p < R / 2; ++p) {
        const int lo = ((p >> tb) << (tb + 1)) | (p & ((1 << tb) - 1)), hi = lo | (1 << tb);
        if ((lo & creg) != creg) continue;
        const Amp x = a[lo], y = a[hi];
        const double* m = o.m;
        switch (o.kind) {
            case OP_GEN:
                a[lo].x = m[0] * x.x - m[1] * x.y + m[2] * y.x - m[3] * y.y;
                a[lo].y = m[0] * x.y + m[1] * x.x + m[2] * y.y + m[3] * y.x;
                a[hi].x = m[4] * x.x - m[5] * x.y + m[6] * y.x - m[7] * y.y;
                a[hi].y = m[4] * x.y + m[5] * x.x + m[6] * y.y + m[7] * y.x;
                break;
            case OP_REAL:
                a[lo].x = m[0] * x.x + m[2] * y.x; a[lo].y = m[0] * x.y + m[2] * y.y;
                a[hi].x = m[4] * x.x + m[6] * y.x; a[hi].y = m[4] * x.y + m[6] * y.y;
                break;
            case OP_RXL:
                a[lo].x = m[0] * x.x - m[3] * y.y; a[lo].y = m[0] * x.y + m[3] * y.x;
                a[hi].x = m[6] * y.x - m[5] * x.y; a[hi].y = m[6] * y.y + m[5] * x.x;
                break;
            case OP_SWAP: a[lo] = y; a[hi] = x; break;
            case OP_YL: a[lo] = {y.y, -y.x}; a[hi] = {-x.y, x.x}; break;
            case OP_DIAG_R:
                if (!(o.flags & 1u)) { a[lo].x = m[0] * x.x - m[1] * x.y; a[lo].y = m[0] * x.y + m[1] * x.x; }
                a[hi].x = m[6] * y.x - m[7] * y.y; a[hi].y = m[6] * y.y + m[7] * y.x;
                break;
            default: break;
        }
    }
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
        uint64_t base = 0;
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
                for (int op = rd.op_begin; op < rd.op_end; ++op) applyOp(a, ops[op], phys);
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
