// Circuit-level peephole pass (see peephole.cpp): gate patterns whose product is diagonal become diagonal gates.
#pragma once
#include <vector>

#include "gate.h"

namespace hyquas {

struct PeepholeStats {
    int gatesIn = 0, gatesOut = 0;
    int zzPatterns = 0;     // cx a,b ; D b ; cx a,b
    int hcxhPatterns = 0;   // h t ; cx *,t ... ; h t
    int xdxPatterns = 0;    // x a ; diagonal gates ; x a
    int mergedPairs = 0;    // adjacent single-qubit gates on one qubit multiplied together
};

std::vector<Gate> peephole(const std::vector<Gate>& gates, PeepholeStats* stats = nullptr);

}  // namespace hyquas
