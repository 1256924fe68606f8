// Data structures shared by the gate-group kernel (group_kernel.cu), its host-side planner and the
// test-only plan emulator (plan_emulator.cpp).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <vector_types.h>

namespace hq {

#ifndef HQ_RBITS
#define HQ_RBITS 4
#endif
constexpr int RBITS = HQ_RBITS;   // register qubits per round (3 or 4; build-time choice)
constexpr int R = 1 << RBITS;     // amplitudes per thread
constexpr int MAX_SEG = 24;
constexpr int MIN_RUN_BITS = 3;   // tiles are made of >= 128-byte contiguous runs

// Arithmetic classes.  Classes < OPK_TEMPLATED are instantiated per (target register bit, control case).
enum OpKind : uint32_t {
    OP_GEN = 0,     // general complex 2x2 on a register bit
    OP_REAL,        // real 2x2 (RY, ...)
    OP_RXL,         // [[a, i b],[i c, d]] with real a,b,c,d (RX, ...)
    OP_SWAP,        // X / CNOT / CCX
    OP_YL,          // [[0, -i],[i, 0]]
    OP_DIAG_R,      // diag(d0, d1), target is a register bit
    OP_ZFLIP,       // diag(1, -1) on a register bit: sign flips only (Z / CZ)
    OP_DIAG_R1,     // diag(1, d1) on a register bit: only the "hi" half is multiplied
    // Butterflies: uncontrolled M = alpha * [[1, p],[q, -p q]] with p, q both in {+1,-1} or both in {+i,-i} (H, RX(+-pi/2),
    // RY(+-pi/2), ...): two FP64 adds per amplitude, no multiplies; alpha is deferred to ONE scalar factor per launch.
    OP_BF0,         // variant v = OP_BFv - OP_BF0 indexes HQ_BF_PQ below
    OP_BF7 = OP_BF0 + 7,
    OPK_TEMPLATED,
    OP_DIAG_T = OPK_TEMPLATED,   // diag(d0, d1) whose target is a thread/outside bit, with register-bit controls
    OP_DIAG_RUN,                 // header of `aux` following entries: diagonal gates that touch no register bit
};
// (p, q) of butterfly variant v as {re p, im p, re q, im q}
constexpr double HQ_BF_PQ[8][4] = {{1, 0, 1, 0},  {1, 0, -1, 0},  {-1, 0, 1, 0},  {-1, 0, -1, 0},
                                   {0, 1, 0, 1},  {0, 1, 0, -1},  {0, -1, 0, 1},  {0, -1, 0, -1}};
// control case of a templated op: 0 = no register-bit control, 1..4 = exactly register bit (cbc-1), 5 = generic mask
constexpr uint32_t CBC_GENERIC = 5;
#define HQ_OP_CODE(kind, tb, cbc) ((uint32_t)(kind) * 24u + (uint32_t)(tb) * 6u + (uint32_t)(cbc))
inline uint32_t op_code(uint32_t kind, uint32_t tb, uint32_t cbc) { return HQ_OP_CODE(kind, tb, cbc); }
constexpr uint32_t CODE_DIAG_T = OPK_TEMPLATED * 24;
constexpr uint32_t CODE_DIAG_RUN = CODE_DIAG_T + 1;

struct alignas(16) DevOp {
    double m[8];       // row-major 2x2 (re, im); diagonal ops use m[0..1] = d0, m[6..7] = d1
    uint32_t code;     // op_code(kind, tbit, cbc) / CODE_DIAG_T / CODE_DIAG_RUN
    uint32_t creg;     // all register-index control bits
    uint32_t flags;    // bit0: d0 == 1 (the "lo" half of a diagonal is untouched); bit2: "special" = has predicate masks, or is a DIAG_T / DIAG_RUN
                       // (top-level ops); bits 16..23: body index of hq_apply_op
    uint32_t aux;      // CODE_DIAG_RUN: number of entries that follow
    uint64_t cphys;    // controls outside the registers, as a mask over the physical local index
    uint64_t tphys;    // diagonal ops with a non-register target: its physical bit (0 = scalar, always d1)
};
static_assert(sizeof(DevOp) == 96, "DevOp layout");

struct alignas(16) DevRound {
    uint16_t ro_in[R];    // shared-memory amplitude index contributed by register index i (read layout)
    uint16_t ro_out[R];   // same for the write layout
    uint64_t go[R];       // physical offset contributed by register index i
    int32_t op_begin, op_end;
    uint32_t flags;       // bit0: read layout != write layout (extra barrier), bit1: last round -> HBM,
                          // bit2: the exchange after this round stays inside each warp (__syncwarp), bit3: so does bit0's
    uint32_t pad;
};

struct GroupParams {
    double2* state;
    uint64_t ntiles;
    uint64_t fixed_base;       // value of the fixed (non-varying) physical bits of this launch
    const uint64_t* run_off;   // [nruns] physical offset of each contiguous run inside a tile
    const DevRound* rounds;
    const DevOp* ops;
    const uint16_t* tb;        // [nrounds][2][NT] shared-memory index of the thread (read, write layout)
    const uint64_t* gt;        // [nrounds][NT] physical offset of the thread (register bits zero)
    int32_t nruns;
    uint32_t run_bytes;
    int32_t nrounds;
    int32_t nops;
    int32_t nseg;
    uint8_t seg_shift[MAX_SEG];  // tile number -> tile base: base |= ((t >> seg_src) & seg_mask) << seg_shift
    uint8_t seg_src[MAX_SEG];
    uint64_t seg_mask[MAX_SEG];
};

}  // namespace hq

struct hq_group_plan {
    int L = 0, K = 0, NT = 0;
    uint64_t tile_mask = 0;
    int nrounds = 0, nops = 0, ngates = 0;
    int nlocal = 0;                       // exchanges between rounds that need only a warp-level barrier
    mutable int grid = 0;
    size_t smem = 0;
    std::vector<unsigned char> blob;      // host image of the device tables (run_off | rounds | ops | gt | tb)
    size_t o_run = 0, o_rounds = 0, o_ops = 0, o_gt = 0, o_tb = 0;
    void* dev_blob = nullptr;             // uploaded at creation when a GPU is bound
    hq::GroupParams p{};                  // pointers refer to dev_blob
    // what the JIT emitter (group_jit.cpp) needs beyond the tables: per round the register qubits and the thread-id-bit ->
    // tile-bit map the tables were built from, and the launch geometry
    struct RoundMeta { int reg[hq::RBITS]; std::vector<int> tbits; };
    std::vector<RoundMeta> meta;
    uint64_t fixed_mask = 0, fixed_value = 0;
    std::string jit_identity;             // what determines this plan's specialised kernel (cache key); empty: interpreter only
    mutable void* jit = nullptr;          // hq::JitKernel*, resolved at the first launch (or by hq_group_plans_warm)
    mutable bool jit_failed = false;
    mutable int jit_occupancy = 0;
    // zero-input variant (the launch's input is |0...0>: no HBM read, no prior zero fill), requested with hq_group_plan_enable_zero_input
    std::string jit_identity_zero;
    mutable void* jit_zero = nullptr;
    mutable bool jit_zero_failed = false;
};
