// Data structures shared by the gate-group kernel (group_kernel.cu), its host-side planner and the
// test-only plan emulator (plan_emulator.cpp).
#pragma once
#include <cstdint>
#include <vector>
#include <vector_types.h>

namespace hq {

constexpr int RBITS = 4;          // register qubits per round
constexpr int R = 1 << RBITS;     // amplitudes per consumer thread
constexpr int NBUF = 3;           // TMA ring depth
constexpr int MAX_SEG = 24;
constexpr int MIN_RUN_BITS = 3;   // tiles are made of >= 128-byte contiguous runs

enum OpKind : uint32_t {
    OP_GEN = 0,     // general complex 2x2 on a register bit
    OP_REAL,        // real 2x2 (H, RY, ...)
    OP_RXL,         // [[a, i b],[i c, d]] with real a,b,c,d (RX, ...)
    OP_SWAP,        // X / CNOT / CCX
    OP_YL,          // [[0, -i],[i, 0]]
    OP_DIAG_R,      // diag(d0, d1), target is a register bit
    OP_DIAG_T,      // diag(d0, d1), target is a thread/outside bit (or none: scalar)
};

struct alignas(16) DevOp {
    double m[8];
    uint32_t kind;
    uint32_t tbit;     // register-index bit of the target (OP_DIAG_T: unused)
    uint32_t creg;     // controls that are register-index bits
    uint32_t flags;    // bit0: d0 == 1 (skip the lo half of a diagonal)
    uint64_t cphys;    // controls outside the registers, as a mask over the physical local index
    uint64_t tphys;    // OP_DIAG_T: physical bit of the target (0 = scalar, always d1)
};
static_assert(sizeof(DevOp) == 96, "DevOp layout");

struct alignas(16) DevRound {
    uint16_t ro_in[R];    // shared-memory amplitude index contributed by register index i (read layout)
    uint16_t ro_out[R];   // same for the write layout
    uint64_t go[R];       // physical offset contributed by register index i
    int32_t op_begin, op_end;
    uint32_t flags;       // bit0: read layout != write layout (extra barrier), bit1: last round -> HBM
    uint32_t pad;
};

struct GroupParams {
    double2* state;
    uint64_t ntiles;
    const uint64_t* run_off;   // [nruns] physical offset of each contiguous run inside a tile
    const DevRound* rounds;
    const DevOp* ops;
    const uint16_t* tb;        // [nrounds][2][NT] shared-memory index of the thread (read, write layout)
    const uint64_t* gt;        // [nrounds][NT] physical offset of the thread (register bits zero)
    int32_t nruns;
    uint32_t run_bytes;
    int32_t nrounds;
    int32_t nseg;
    uint8_t seg_shift[MAX_SEG];  // tile number -> tile base: base |= ((t >> seg_src) & seg_mask) << seg_shift
    uint8_t seg_src[MAX_SEG];
    uint64_t seg_mask[MAX_SEG];
};

}  // namespace hq

struct hq_group_plan {
    int L = 0, K = 0, NT = 0;
    uint64_t tile_mask = 0;
    int nrounds = 0, nops = 0, grid = 0;
    size_t smem = 0;
    std::vector<unsigned char> blob;      // host image of the device tables (run_off | rounds | ops | gt | tb)
    size_t o_run = 0, o_rounds = 0, o_ops = 0, o_gt = 0, o_tb = 0;
    void* dev_blob = nullptr;             // uploaded at creation when a GPU is bound
    hq::GroupParams p{};                  // pointers refer to dev_blob
};
