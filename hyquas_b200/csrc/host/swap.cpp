#include "swap.h"

#include "circuit.h"

namespace hyquas {

static void notBuilt() {
    fprintf(stderr, "multi-GPU swap layer is not available in this build\n");
    exit(1);
}
void commInitFromEnv() { notBuilt(); }
void bcastAmp(qComplex*, int) { notBuilt(); }
void gatherItems(std::vector<ResultItem>&) { notBuilt(); }
SwapExec::SwapExec(qComplex* s, int L, const SwapPlan& p) : state(s), numLocal(L), plan(p) {}
void SwapExec::begin() { notBuilt(); }
int SwapExec::waitNextChunk() { notBuilt(); return 0; }
void SwapExec::end() { notBuilt(); }

}  // namespace hyquas
