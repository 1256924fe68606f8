// User-facing circuit object: the same surface as the reference's Circuit / ResultItem (src/circuit.h:9-47):
//   Circuit c(n); c.addGate(Gate::H(0)); c.compile(); int us = c.run(); c.printState();
// run() returns the elapsed microseconds of the execution phase ("Time Cost", src/circuit.cpp:22,46-52).
#pragma once

#include <string>
#include <vector>
#include "utils.h"
#include "gate.h"
#include "schedule.h"

struct ResultItem {   // one line of the amplitude dump: logical index + amplitude
    qindex idx;
    qComplex amp;
    ResultItem() = default;
    ResultItem(const qindex& i, const qComplex& a): idx(i), amp(a) {}
    std::string str() const;
    void print() { fputs(str().c_str(), stdout); }
    bool operator < (const ResultItem& o) const { return idx < o.idx; }
};

class Circuit {
public:
    const int numQubits;
    Circuit(int n): numQubits(n) {}
    ~Circuit();

    // ---- the reference's surface --------------------------------------------------------------------------
    void addGate(const Gate& g) { gates.push_back(g); }
    void compile();                                          // peephole -> partition -> device plans
    int run(bool copy_back = true, bool destroy = true);     // -> microseconds of the execution phase
    void printState();
    void dumpGates();
    ResultItem ampAt(qindex idx);
    qComplex ampAtGPU(qindex idx);

    // ---- extras used by the C-ABI / tests / bench -----------------------------------------------------------
    void prepareState();                                     // (re)allocate + |0..0>
    int execute(std::vector<float>* perGroupMs = nullptr, bool stateIsGarbage = false);   // the timed part of run() on the resident state
    void allocState();                                       // allocate (and map to the peers) without initialising
    void destroyState();
    std::string stateDump();                                 // the text printState() prints
    bool fullState(std::vector<qComplex>& out);              // all 2^n amplitudes in LOGICAL order (single process, small n)
    bool localShard(double* out);                            // this process' amplitudes in PHYSICAL order
    double measure(int logicalQubit);                        // P(qubit reads 0) over the distributed state (collective)
    double norm2();                                          // sum |a|^2 over this process' shard
    double swapAloneMs();                                    // the schedule's exchanges with no compute (collective)
    std::string compileError() const;                        // why compile() would refuse (empty: it would not)
    size_t planBytes() const;                                // bytes of device tables uploaded by compile()
    size_t dumpBytes() const { return dumpItems.size() * sizeof(ResultItem); }
    const Schedule& getSchedule() const { return schedule; }
    const std::vector<Gate>& getGates() const { return gates; }
    double lastDeviceMs = 0;                                 // CUDA-event time of the last run()

private:
    qindex toPhysicalID(qindex idx);
    qindex toLogicID(qindex idx);
    void collectDump();
    std::vector<Gate> gates;
    std::vector<qComplex*> deviceStateVec;
    Schedule schedule;
    std::vector<qComplex> result;                            // host copy in PHYSICAL order (copy_back)
    std::vector<ResultItem> dumpItems;                       // what printState() shows, captured before destroy
    bool compiled = false;
};
