// Time predictor used by the partitioner to price a gate group on this GPU.
// Same role and entry points as the reference's Evaluator singleton (src/evaluator.h:16-149,
// src/evaluator.cpp:111-229) but the model is a roofline: a group costs
//     max( sweep time of the local state at measured HBM bandwidth,  sum of per-gate in-register cost )
// with per-gate-class costs calibrated by the B200 microbenchmarks (tools/calibrate.py writes the parameter
// file; built-in defaults are the values measured on this pool's B200).
#pragma once
#include <set>
#include <string>
#include <vector>

#include "gate.h"
#include "schedule.h"
#include "utils.h"

class Evaluator {
public:
    static Evaluator* getInstance();
    // predicted milliseconds for one gate-group launch over 2^numQubits local amplitudes
    double perfPerGate(int numQubits, const GateGroup* gg);
    double perfPerGate(int numQubits, const std::vector<GateType>& types);
    double perfPerGate(int numQubits, const std::vector<Gate>& gates);
    // predicted milliseconds for one fused dense (TransMM) launch with a 2^blasSize x 2^blasSize matrix
    double perfBLAS(int numQubits, int blasSize);
    // one fused dense launch applying several matrices (qubit counts in `ms`) in a single sweep
    double perfDense(int numQubits, const std::vector<int>& ms);
    double sweepMs30() const { return 32.0 * 1073741824.0 / (hbmGBs * 1e9) * 1e3; }
    // predicted milliseconds for swapping k local bits with k global bits (NVLink bytes / measured bandwidth)
    double perfSwap(int numQubits, int k);
    double nvlinkGBs;                       // per-direction bandwidth of one GPU during the exchange
    bool PerGateOrBLAS(const GateGroup* gg_pergate, const GateGroup* gg_blas, int numQubits, int blasSize);
    void loadParam(int numQubits);          // optional overrides: {L}qubits.out in the reference's layout (hq_preprocess), $HYQUAS_PARAM_FILE
    bool loadReferenceLayout(int numQubits);
    // model constants (public so that the calibration tool and tests can read/write them)
    double hbmGBs;                          // achieved sweep bandwidth of the gate-group kernel (read+write)
    double gateNs[32];                      // per-gate cost per 2^30 amplitudes in ms, indexed by GateType
    double groupBaseMs30;                   // fixed part of a gate-group launch that does not overlap the sweep
    double denseBaseMs30;                   // same for the fused dense kernel
    double roundMs30;                       // cost of one extra register round per 2^30 amplitudes, ms
    double circuitFactor;                   // in-circuit / microbenchmark cost ratio of the tile kernel's gate work (fit)
    double denseMs30[8];                    // fused dense kernel, per 2^30 amplitudes, indexed by matrix qubits
    double launchMs;                        // fixed per-launch overhead
    // specialised (JIT) tile kernels: a gate costs the FP64 instructions per amplitude its matrix needs (0 for permutations and
    // +-1 / +-i phases, 2 for butterflies and real / RX-like 2x2, 6 for a general complex 2x2, 1 per multiplied half for a phase)
    bool specialised;                       // price tile groups for the specialised kernels (hq_jit_available)
    double instrMs30;                       // ms per (FP64 instruction per amplitude) over 2^30 amplitudes
    double jitRoundMs30;                    // coefficient flush + shared-memory exchange of one extra round
    double jitBaseMs30;
    double jitUnderSweepMs30;               // what an instruction per amplitude costs a launch that is sweep-bound
    static double instrPerAmp(const Gate& g);
    static double instrPerAmpUncached(const Gate& g);
    static double fusedInstr(const std::vector<Gate>& gates);    // instruction count if every one- / two-qubit block fuses
    bool fusionAware;                       // price tile groups with block fusion in mind (HQ_EVAL_FUSION=0 switches it off)
    static int registerRounds(const std::vector<Gate>& gates);   // rounds the tile kernel will need for this group
    unsigned long long signature(int numQubits);                 // hash of every model constant in force for this size
private:
    Evaluator();
    bool loaded = false;
    std::set<int> triedLayout;
};
