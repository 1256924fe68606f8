/*
 * hyquas_b200.h -- C-ABI of libhyquas_b200.so (sm_100a).
 *
 * This is the drop-in boundary: plain pointers, sizes and POD structs only, every entry point
 * returns an int status (0 = HQ_OK).  Each block below names the reference interface it stands in
 * for (paths relative to the HyQuas tree).  The reference's public C++ surface (Circuit / Gate /
 * MyGlobalVars / Logger, src/circuit.h, src/gate.h, src/utils.h, src/logger.h) is re-implemented on top of
 * this layer in hyquas_b200/csrc/host/ with the same names, so main.cpp and the micro-benchmark drivers of the
 * reference compile unchanged against it; see INTEGRATION.md.
 */
#ifndef HYQUAS_B200_H
#define HYQUAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HQ_OK 0
#define HQ_ERR_CUDA 1       /* a CUDA runtime call failed; message via hq_last_error() */
#define HQ_ERR_ARG 2        /* invalid argument (bad mask, target outside the tile, ...) */
#define HQ_ERR_UNSUPPORTED 3
#define HQ_ERR_NCCL 4

/* Gate type numbering == enum class GateType of the reference (src/gate.h:7-9); the evaluator
 * preprocess tool iterates these by integer value (evaluator-preprocess/process.cpp:26,55). */
enum hq_gate_type {
    HQ_CCX = 0, HQ_CNOT, HQ_CY, HQ_CZ, HQ_CRX, HQ_CRY, HQ_CU1, HQ_CRZ, HQ_U1, HQ_U2, HQ_U3, HQ_H, HQ_X, HQ_Y, HQ_Z,
    HQ_S, HQ_SDG, HQ_T, HQ_TDG, HQ_RX, HQ_RY, HQ_RZ, HQ_TOTAL, HQ_ID, HQ_GII, HQ_GZZ, HQ_GOC, HQ_GCC
};

/* One lowered gate of a group, addressed by PHYSICAL LOCAL bit positions of the amplitude index
 * (0 .. L-1).  Replaces KernelGate (src/gate.h:68-110): there, operands are shared-memory bit
 * indices or block-index bit indices plus an *IsGlobal flag; here the kernel derives that itself
 * from tile_mask.  target == -1 means "no target" (a scalar applied to every amplitude: GII/GZZ/GCC
 * produced when a diagonal gate sits on a global qubit, src/executor.cpp:286-398).
 * mat is row-major {m00, m01, m10, m11} x {re, im}, i.e. KernelGate's r00,i00,...,r11,i11. */
typedef struct hq_gate {
    int32_t type;      /* enum hq_gate_type; only used to pick specialised arithmetic, mat is authoritative */
    int32_t target;    /* physical local bit, or -1 */
    int32_t control;   /* physical local bit, or -1 */
    int32_t control2;  /* physical local bit, or -1 */
    double  mat[8];
} hq_gate;

const char* hq_last_error(void);
const char* hq_version(void);

/* ---- runtime (replaces MyGlobalVars::init, src/utils.cpp:17-60) ------------------------------- */
int hq_device_count(int* n);
int hq_init(int device);                       /* bind this process to one GPU; creates the compute + comm streams */
int hq_shutdown(void);
int hq_sync(void);                             /* Executor::finalize's cudaStreamSynchronize, src/executor.cpp:634-640 */
int hq_device_info(char* name, size_t cap, int* sm_count, size_t* total_mem);

/* ---- state vector (replaces kernelInit / kernelDeviceToHost / kernelGetAmp / kernelDestroy,
 *      src/kernelSimple.cu:9-37,518-530).  One allocation of 16 * 2^L bytes: every kernel here is
 *      in place, so the reference's doubled buffer (kernelSimple.cu:10-13) does not exist. ------- */
int hq_state_alloc(int L, void** state);
int hq_state_free(void* state);
int hq_state_init(void* state, int L, int set_amp0);          /* zero; amp[0] = 1 if set_amp0 (rank holding |0..0>) */
int hq_state_download(const void* state, int L, int64_t first, int64_t count, double* host_re_im);
int hq_state_upload(void* state, int L, int64_t first, int64_t count, const double* host_re_im);
int hq_amp_fetch(const void* state, int64_t idx, double out_re_im[2]);
/* On-device threshold scan used by printState (src/circuit.cpp:299-309) so that 32+ qubit states
 * never travel to the host: returns physical local indices (ascending) with |a|^2 > thresh. */
int hq_dump_scan(const void* state, int L, double thresh, int64_t* idx_out, double* amp_out, int64_t cap, int64_t* found);
int hq_state_norm2(const void* state, int L, double* out);
/* kernelMeasure (src/kernel.h:14, src/kernelSimple.cu:482-516): probability that physical local bit target_bit reads 0 */
int hq_state_measure(const void* state, int L, int target_bit, double* p0);

/* ---- gate-group kernel (replaces copyGatesToSymbol + launchExecutor, src/kernel.h:25-28,
 *      src/kernelOpt.cu:425-433,499-506, and the mask bookkeeping of Executor::prepareBitMap /
 *      getLogicShareMap, src/executor.cpp:598-632).
 *      tile_mask selects the physical bits that vary inside one tile (popcount = hq_group_tile_bits(),
 *      low hq_group_min_run_bits() bits must be set); every non-diagonal target must lie in the tile;
 *      controls and diagonal targets may be anywhere in [0, L).  state is updated in place. ---------- */
typedef struct hq_group_plan hq_group_plan;
int hq_group_tile_bits(void);
int hq_group_min_run_bits(void);
int hq_group_plan_create(int L, uint64_t tile_mask, const hq_gate* gates, int ngates, hq_group_plan** plan);
/* _ex: the launch covers only the amplitudes whose fixed_mask bits equal fixed_value (one chunk of the state while the
 * other chunks are still on the wire, src/executor.cpp:495-531 applyPerGateGroupSliced); gates must not touch fixed bits. */
int hq_group_plan_create_ex(int L, uint64_t tile_mask, uint64_t fixed_mask, uint64_t fixed_value, const hq_gate* gates, int ngates,
                            hq_group_plan** plan);
int hq_group_plan_launch(const hq_group_plan* plan, void* state, int on_comm_stream);
int hq_group_plan_info(const hq_group_plan* plan, int* rounds, int* ops, int* grid, int* smem_bytes);
int hq_group_plan_table_bytes(const hq_group_plan* plan, int* bytes);   /* size of the uploaded device tables */
int hq_group_plan_local_exchanges(const hq_group_plan* plan, int* n);   /* round exchanges done with a warp-level barrier */
int hq_group_plan_destroy(hq_group_plan* plan);
/* Gate groups run as per-group specialised kernels (NVRTC, sm_100a; cached in memory and under $HQ_JIT_CACHE or
 * ~/.cache/hyquas_b200/jit; HQ_JIT=0 keeps the interpreter kernel).  A plan compiles at its first launch; warming a whole
 * schedule's plans at once compiles the cache misses on all host cores. */
/* what the specialised kernel of this plan costs: register rounds and FP64 instructions per amplitude (works host-only) */
int hq_group_plan_cost(const hq_group_plan* plan, int* rounds, double* fp64_per_amp);
/* The first group of a circuit acts on |0...0>: its zero-input variant reads nothing and needs no zero-filled state (kernelInit's
 * memset + the first sweep's read, src/kernelSimple.cu:9-37).  enable before hq_group_plans_warm; launch_from_zero returns
 * HQ_ERR_UNSUPPORTED when the variant is unavailable (zero-fill and launch normally then). */
int hq_group_plan_enable_zero_input(hq_group_plan* plan);
int hq_group_plan_launch_from_zero(const hq_group_plan* plan, void* state, int has_amp0);
int hq_group_plans_warm(hq_group_plan* const* plans, int n);
int hq_group_plan_is_specialised(const hq_group_plan* plan, int* yes);
int hq_jit_available(int* yes);   /* HQ_JIT not 0 and NVRTC loadable; the evaluator prices tile groups accordingly */
int hq_jit_stats(int* kernels_loaded, int* compiled, int* disk_hits, double* compile_seconds);
/* directory of the on-disk caches (compiled kernels, partitioner search results): $HQ_JIT_CACHE, else
 * $XDG_CACHE_HOME/hyquas_b200/jit, else ~/.cache/hyquas_b200/jit; "" when HQ_JIT_CACHE=off */
int hq_cache_dir(char* out, size_t cap);
int hq_group_apply(void* state, int L, uint64_t tile_mask, const hq_gate* gates, int ngates);   /* create+launch+destroy */

/* ---- fused dense-matrix kernel (replaces the TransMM path: cuttExecute + cublasZgemm in Executor::applyBlasGroup,
 *      src/executor.cpp:533-575, and the device upload of GateGroup::initGPUMatrix, src/schedule.cpp:704-721).
 *      One in-place sweep applies nmat dense matrices in order; matrix i acts on m_list[i] qubits whose physical
 *      local bit positions are the next m_list[i] entries of qubit_pos (entry b = bit b of the row/column index);
 *      U is column-major, interleaved (re, im), 2^m x 2^m, like the A operand of the reference's Zgemm call.
 *      All matrices of one plan must fit in one tile: their qubits together with physical bits 0..2 span <= 12 bits. */
typedef struct hq_dense_plan hq_dense_plan;
int hq_dense_plan_create(int L, int nmat, const int* m_list, const int* qubit_pos, const double* u_colmajor, hq_dense_plan** plan);
int hq_dense_plan_create_ex(int L, uint64_t fixed_mask, uint64_t fixed_value, int nmat, const int* m_list, const int* qubit_pos,
                            const double* u_colmajor, hq_dense_plan** plan);
int hq_dense_plan_launch(const hq_dense_plan* plan, void* state, int on_comm_stream);
int hq_dense_plan_info(const hq_dense_plan* plan, int* tile_bits, int* smem_bytes, int* grid, double* flops_per_amp, int* table_bytes);
int hq_dense_plan_destroy(hq_dense_plan* plan);
int hq_dense_apply(void* state, int L, int m, const int* qubit_pos, const double* u_colmajor);   /* create+launch+destroy */

/* ---- multi-GPU: one process per GPU (replaces the NCCL bootstrap of MyGlobalVars::init, src/utils.cpp:46-58, and
 *      Executor::transpose + all2all + sliceBarrier, src/executor.cpp:59-179,650-659).
 *      A swap trades k local bits with k global bits IN PLACE: the local state is 2^k chunks (the values of the k swapped local
 *      bits); chunk c goes to the rank whose swapped global bits equal c and is replaced by that rank's chunk.  Two transports,
 *      chosen at hq_comm_init (HQ_SWAP=p2p|nccl; p2p whenever every GPU pair is peer-capable):
 *        p2p   the state allocations are mapped into every process (CUDA IPC, hq_swap_attach); one kernel per exchange step swaps
 *              the two chunks of a rank pair element by element over NVLink.  No staging, no second buffer, and the swapped
 *              local bits may be ANY positions >= 3 (hq_swap_any_position = 1): no local bit permutation is needed first.
 *        nccl  chunks move in pieces through a two-slot staging ring (ncclSend/ncclRecv on the comm stream + un-stage copies);
 *              the swapped bits must be the TOP k local positions (hq_state_bitswap brings them there).
 *      Either way one event per chunk: hq_swap_wait_chunk() orders the compute stream behind one chunk at a time, so per-chunk
 *      gate groups (plans created with fixed bits = the swapped positions) run under the rest of the exchange. ---------------- */
typedef struct hq_swap_plan hq_swap_plan;
int hq_comm_unique_id(unsigned char out[128]);                       /* ncclGetUniqueId (rank 0), to be broadcast by the host */
int hq_comm_init(int world, int rank, const unsigned char id[128]);  /* ncclCommInitRank */
int hq_comm_info(int* world, int* rank);
int hq_comm_destroy(void);
int hq_comm_bcast_host(void* buf, size_t bytes, int root);           /* control plane of printState (small host buffers) */
int hq_comm_allgather_host(const void* send, void* recv, size_t bytes_per_rank);
int hq_swap_any_position(int* any);      /* 1: p2p transport, swapped local bits may be any positions >= 3; 0: top k only */
int hq_swap_attach(void* state);         /* p2p: map every rank's state into this process (collective, untimed) */
int hq_swap_detach(void);
int hq_state_bitswap(void* state, int L, int npairs, const int* a, const int* b);   /* in-place local bit permutation */
int hq_swap_plan_create(int L, int k, const int* local_bits, const int* global_bits, hq_swap_plan** plan);
int hq_swap_plan_set_overlap(hq_swap_plan* plan, int groups_under_exchange);   /* picks the exchange kernel's shape */
int hq_swap_begin(hq_swap_plan* plan, void* state);                  /* enqueue the whole exchange on the comm stream */
int hq_swap_wait_chunk(hq_swap_plan* plan, int* chunk);              /* compute stream waits for the next landed chunk */
int hq_swap_end(hq_swap_plan* plan);
int hq_swap_plan_destroy(hq_swap_plan* plan);

/* ---- timing helpers (cudaEvent pairs on the compute stream; MEASURE_STAGE, src/executor.cpp:406-458) */
int hq_timer_start(void);
int hq_timer_stop_ms(float* ms);

/* ---- B200 microbenchmarks feeding the evaluator (replaces evaluator-preprocess/process.cpp:85-165, which times
 *      cublasZgemm and cuTT; here: FP64 FMA rate (kind 0), FP64 tensor-core mma.sync rate (kind 1), copy bandwidth) */
int hq_microbench_fp64(int kind, double* tflops);
int hq_microbench_copy(void* state, int L, double* gbs);

#ifdef __cplusplus
}
#endif
#endif /* HYQUAS_B200_H */
