/*
 * hyquas_b200_circuit.h -- circuit-level C-ABI of libhyquas_b200.so.
 *
 * A thin extern "C" shell around the C++ surface that mirrors the reference's Circuit API
 * (src/circuit.h:20-34: Circuit(n), addGate, compile, run, printState) so that ctypes / cgo / JNI style
 * hosts can drive it.  Status codes as in hyquas_b200.h; messages via hq_circuit_last_error().
 */
#ifndef HYQUAS_B200_CIRCUIT_H
#define HYQUAS_B200_CIRCUIT_H

#include <stddef.h>
#include "hyquas_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hq_circuit hq_circuit;

const char* hq_circuit_last_error(void);
int hq_runtime_init(void);                                   /* MyGlobalVars::init(): bind GPU, (multi-process) NCCL */
int hq_runtime_init_host_only(int world_size, int rank);     /* partitioner/plan tests without a GPU */

int hq_circuit_create(int num_qubits, hq_circuit** out);     /* Circuit(int numQubits)            circuit.h:22 */
int hq_circuit_from_qasm(const char* text, hq_circuit** out);/* parse_circuit                     main.cpp:68-231 */
/* Gate::<type>(...) + Circuit::addGate; type = enum hq_gate_type, unused operands -1     gate.h:30-57, circuit.h:25 */
int hq_circuit_add_gate(hq_circuit* c, int type, int control2, int control, int target, const double* params, int nparams);
int hq_circuit_num_qubits(const hq_circuit* c);
int hq_circuit_num_gates(const hq_circuit* c);
int hq_circuit_compile(hq_circuit* c);                       /* Circuit::compile                  circuit.cpp:177-210 */
int hq_circuit_plan_only(hq_circuit* c, int* stages, int* groups, int* swapped_bits);   /* Compiler::run only (host) */
/* Circuit::run: returns wall microseconds of the execution phase ("Time Cost") and the CUDA-event time */
int hq_circuit_run(hq_circuit* c, int copy_back, int destroy, int* time_us, double* device_ms);
/* run() split in two so that a resident state can be re-used: allocate + |0..0>, then the timed execution phase.
 * per_group_ms (optional): CUDA-event time of each gate-group launch (MEASURE_STAGE, src/executor.cpp:406-458). */
int hq_circuit_prepare_state(hq_circuit* c);
int hq_circuit_execute(hq_circuit* c, int* time_us, double* device_ms, float* per_group_ms, int cap, int* ngroups);
int hq_circuit_release_state(hq_circuit* c);                  /* free the resident state, keep the compiled schedule */
int hq_circuit_norm2(hq_circuit* c, double* out);
int hq_circuit_measure(hq_circuit* c, int qubit, double* p0);   /* P(logical qubit reads 0); collective (kernelMeasure, src/kernel.h:14) */
/* the schedule's global<->local exchanges alone, back to back (collective; leaves the state permuted) */
int hq_circuit_swap_alone_ms(hq_circuit* c, double* ms);
int hq_circuit_io_bytes(const hq_circuit* c, size_t* h2d_plan_bytes, size_t* d2h_dump_bytes);
int hq_circuit_schedule_info(const hq_circuit* c, int* stages, int* groups, int* gates_in_groups);
int hq_circuit_group_info(const hq_circuit* c, int index, int* backend, int* ngates, double* predicted_ms, int* launches, int* nblocks);
int hq_circuit_group_cost(const hq_circuit* c, int index, int* rounds, double* fp64_per_amp);   /* tile groups: rounds, FP64 instr / amplitude */
int hq_circuit_dump(hq_circuit* c, char* buf, size_t cap, size_t* needed);              /* printState text */
int hq_circuit_amplitudes(hq_circuit* c, double* out_re_im); /* all 2^n amplitudes, logical order (small n) */
int hq_circuit_amp_at(hq_circuit* c, long long idx, double out_re_im[2]);   /* Circuit::ampAt; collective across processes */
int hq_circuit_local_shard(hq_circuit* c, double* out_re_im);  /* this process' 2^(n-g) amplitudes, physical order */
int hq_circuit_final_layout(hq_circuit* c, int* pos);         /* pos[logical qubit] = physical bit (Schedule::finalState) */
int hq_circuit_logger_flush(char* buf, size_t cap);          /* Logger::print */
int hq_circuit_destroy(hq_circuit* c);

#ifdef __cplusplus
}
#endif
#endif
