#include "schedule.h"

std::string to_string(Backend b) {
    switch (b) {
        case Backend::PerGate: return "PerGate";
        case Backend::BLAS: return "BLAS";
        default: return "None";
    }
}

int Schedule::numGroups() const {
    int s = 0;
    for (const auto& lg : localGroups) s += (int)lg.overlapGroups.size() + (int)lg.fullGroups.size();
    return s;
}

// SHOW_SCHEDULE-style listing: one block per gate group, one line per gate.
void Schedule::dump(int numQubits) const {
    int stage = 0;
    for (const auto& lg : localGroups) {
        printf("=== stage %d: local qubits", stage++);
        for (int q = 0; q < numQubits; q++) if (lg.contains(q)) printf(" %d", q);
        printf("\n    pos:");
        for (int q = 0; q < numQubits; q++) printf(" %d", lg.state.pos[q]);
        printf("\n");
        auto show = [&](const GateGroup& gg, const char* kind) {
            printf("<%s %s> %d gates, tile/matrix qubits:", to_string(gg.backend).c_str(), kind, (int)gg.gates.size());
            for (int q = 0; q < numQubits; q++) if (gg.contains(q)) printf(" %d", q);
            printf("\n");
            for (const Gate& g : gg.gates) {
                printf("    %-4s t=%d", g.name.c_str(), g.targetQubit);
                if (g.controlQubit >= 0) printf(" c=%d", g.controlQubit);
                if (g.controlQubit2 >= 0) printf(" c2=%d", g.controlQubit2);
                printf("\n");
            }
        };
        for (const auto& gg : lg.overlapGroups) show(gg, "overlap");
        for (const auto& gg : lg.fullGroups) show(gg, "full");
    }
    fflush(stdout);
}
