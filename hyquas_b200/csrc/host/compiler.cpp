#include "compiler.h"

#include <algorithm>
#include <cassert>
#include <unistd.h>

#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>

#include "evaluator.h"
#include "logger.h"

namespace hyquas {

// A gate acts "non-diagonally" on its target unless its matrix is diagonal, and "diagonally" on its controls.
// Two gates commute when every qubit they share is acted on diagonally by both.  A skipped gate therefore
// blocks (a) everything on its non-diagonal qubit, (b) non-diagonal action on its diagonal qubits.
std::vector<int> runnableGates(const std::vector<Gate>& gates, const std::vector<int>& order, qindex tileSet, int cap) {
    std::vector<int> out;
    qindex blockedAll = 0, blockedNonDiag = 0;
    int seen = 0;
    for (int gi : order) {
        if (++seen > cap) break;
        const Gate& g = gates[gi];
        const bool diag = g.isDiagonal();
        qindex qn = 0, qd = 0;
        (diag ? qd : qn) |= qindex(1) << g.targetQubit;
        if (g.controlQubit >= 0) qd |= qindex(1) << g.controlQubit;
        if (g.controlQubit2 >= 0) qd |= qindex(1) << g.controlQubit2;
        bool ok = !(qn & (blockedAll | blockedNonDiag)) && !(qd & blockedAll);
        if (ok && !diag && !(tileSet >> g.targetQubit & 1)) ok = false;
        if (ok) out.push_back(gi);
        else { blockedAll |= qn; blockedNonDiag |= qd; }
    }
    return out;
}

// Same scan for a dense block: a gate may join only if ALL its qubits (controls and diagonal targets too) lie in
// `qset` -- the block's matrix acts on nothing else (the reference's enableGlobal=false mode, src/compiler.cpp:265-278).
std::vector<int> runnableDense(const std::vector<Gate>& gates, const std::vector<int>& order, qindex qset, int cap) {
    std::vector<int> out;
    qindex blockedAll = 0, blockedNonDiag = 0;
    int seen = 0;
    for (int gi : order) {
        if (++seen > cap) break;
        const Gate& g = gates[gi];
        const bool diag = g.isDiagonal();
        qindex qn = 0, qd = 0;
        (diag ? qd : qn) |= qindex(1) << g.targetQubit;
        if (g.controlQubit >= 0) qd |= qindex(1) << g.controlQubit;
        if (g.controlQubit2 >= 0) qd |= qindex(1) << g.controlQubit2;
        bool ok = !(qn & (blockedAll | blockedNonDiag)) && !(qd & blockedAll);
        if (ok && ((qn | qd) & ~qset)) ok = false;
        if (ok) out.push_back(gi);
        else { blockedAll |= qn; blockedNonDiag |= qd; }
    }
    return out;
}

// The searches below ask the same question for many candidate qubits at once ("how many gates could run if q joined the set?").
// FrontierScan answers it for all candidates in ONE pass over the gate list: per-gate action masks are computed once per list and
// every candidate keeps its own pair of blocked masks (same rules as runnableGates, which the tests compare it with).
struct FrontierScan {
    std::vector<qindex> qn, qd;   // per gate: its non-diagonal target bit (0 for a diagonal gate) / the qubits it acts on diagonally
    explicit FrontierScan(const std::vector<Gate>& gates) : qn(gates.size()), qd(gates.size()) {
        for (size_t i = 0; i < gates.size(); i++) {
            const Gate& g = gates[i];
            qindex n = 0, d = 0;
            (g.isDiagonal() ? d : n) |= qindex(1) << g.targetQubit;
            if (g.controlQubit >= 0) d |= qindex(1) << g.controlQubit;
            if (g.controlQubit2 >= 0) d |= qindex(1) << g.controlQubit2;
            qn[i] = n; qd[i] = d;
        }
    }
    // count[c] = |runnableGates(order, set | extra[c], cap)|
    void counts(const std::vector<int>& order, qindex set, int cap, const std::vector<qindex>& extra, std::vector<int>& count) const {
        const size_t nc = extra.size();
        std::vector<qindex> bAll(nc, 0), bNd(nc, 0);
        count.assign(nc, 0);
        int seen = 0;
        for (int gi : order) {
            if (++seen > cap) break;
            const qindex n = qn[gi], d = qd[gi], outside = n & ~set;
            for (size_t c = 0; c < nc; c++) {
                const qindex blocked = (n & (bAll[c] | bNd[c])) | (d & bAll[c]) | (outside & ~extra[c]);
                const qindex m = blocked ? ~qindex(0) : 0;
                bAll[c] |= n & m;
                bNd[c] |= d & m;
                count[c] += blocked ? 0 : 1;
            }
        }
    }
    std::vector<int> runnable(const std::vector<int>& order, qindex set, int cap) const {
        std::vector<int> out;
        qindex bAll = 0, bNd = 0;
        int seen = 0;
        for (int gi : order) {
            if (++seen > cap) break;
            const qindex n = qn[gi], d = qd[gi];
            if ((n & (bAll | bNd)) | (d & bAll) | (n & ~set)) { bAll |= n; bNd |= d; }
            else out.push_back(gi);
        }
        return out;
    }
};

// Move to a layout in which exactly `newLocals` are local: every outgoing qubit trades places with one incoming qubit.
// With the p2p transport (anyBit) the outgoing qubit is traded from wherever it sits (positions < 3 excepted: they stay
// in every tile); with the nccl transport outgoing qubits are first brought to the top local positions by in-place bit
// swaps so that the chunks are contiguous.
SwapPlan planSwap(State& state, qindex newLocals, int numQubits, int numLocal, bool anyBit) {
    SwapPlan plan;
    std::vector<int> outgoing, incoming;
    for (int p = 0; p < numLocal; p++) if (!(newLocals >> state.layout[p] & 1)) outgoing.push_back(state.layout[p]);
    for (int p = numLocal; p < numQubits; p++) if (newLocals >> state.layout[p] & 1) incoming.push_back(state.layout[p]);
    assert(outgoing.size() == incoming.size());
    const int k = (int)outgoing.size();
    if (anyBit) {
        const int minBit = 5;   // positions below stay in every tile (pinnedBits) and are never traded
        for (int q : outgoing) {
            if (state.pos[q] >= minBit) continue;
            // rare: an outgoing qubit sits in the always-in-tile low bits; move it to the highest staying position
            for (int p = numLocal - 1; p >= minBit; p--) {
                const int other = state.layout[p];
                if (std::find(outgoing.begin(), outgoing.end(), other) != outgoing.end()) continue;
                plan.localPerm.push_back({state.pos[q], p});
                state.swapPhysical(state.pos[q], p);
                break;
            }
        }
        std::sort(outgoing.begin(), outgoing.end(), [&](int x, int y) { return state.pos[x] < state.pos[y]; });
    } else {
        // outgoing qubits that already sit in the top-k window keep their place
        std::vector<int> slots;
        for (int p = numLocal - k; p < numLocal; p++) slots.push_back(p);
        std::vector<int> movers;
        for (int q : outgoing) {
            auto it = std::find(slots.begin(), slots.end(), state.pos[q]);
            if (it != slots.end()) slots.erase(it); else movers.push_back(q);
        }
        for (size_t i = 0; i < movers.size(); i++) {
            const int a = state.pos[movers[i]], b = slots[i];
            plan.localPerm.push_back({a, b});
            state.swapPhysical(a, b);
        }
        outgoing.clear();
        for (int p = numLocal - k; p < numLocal; p++) outgoing.push_back(state.layout[p]);
    }
    std::sort(incoming.begin(), incoming.end(), [&](int x, int y) { return state.pos[x] < state.pos[y]; });
    for (int i = 0; i < k; i++) {
        const int lb = state.pos[outgoing[i]], gb = state.pos[incoming[i]];
        plan.localBit.push_back(lb);
        plan.globalBit.push_back(gb - numLocal);
        state.swapPhysical(lb, gb);
    }
    return plan;
}

}  // namespace hyquas

Compiler::Compiler(int numQubits_, std::vector<Gate> inputGates)
    : numQubits(numQubits_), numLocal(numQubits_ - MyGlobalVars::bit), gates(std::move(inputGates)) {
    // the partitioner matches gates by id (dense pick, deferral): ids are positions in THIS list, whatever the caller's Gate
    // objects carried (a Gate added twice keeps one gateID in the reference's scheme, src/gate.cpp:7)
    for (size_t i = 0; i < gates.size(); i++) gates[i].gateID = (int)i;
    tileBits = std::min(hq_group_tile_bits(), numLocal);
    pinnedBits = std::min(5, tileBits);   // planSwap relies on pinnedBits <= 5
    if (const char* e = getenv("HQ_PINNED_BITS")) pinnedBits = std::max(hq_group_min_run_bits(), std::min(atoi(e), std::min(5, tileBits)));
    backendMode = 4;   // the reference's BACKEND numbering: 1 = group (tile kernel only), 3 = blas (dense only), 4 = mix
    if (const char* e = getenv("HQ_BACKEND")) {
        const std::string v = e;
        backendMode = (v == "group" || v == "1") ? 1 : ((v == "blas" || v == "3") ? 3 : 4);
    }
    matLimit = 6;
    if (const char* e = getenv("HQ_MAT")) matLimit = std::max(3, std::min(atoi(e), 6));
    overlapSlack = 1.0;
    if (const char* e = getenv("HQ_OVERLAP_SLACK")) overlapSlack = atof(e);
    enableOverlap = true;
    if (const char* e = getenv("HQ_ENABLE_OVERLAP")) enableOverlap = atoi(e) != 0;
    if (const char* e = getenv("HQ_OVERLAP_MODE")) {
        const std::string v = e;
        if (v == "off") enableOverlap = false;
    }
    rebalanceGroups = true;
    if (const char* e = getenv("HQ_REBALANCE")) rebalanceGroups = atoi(e) != 0;
    cutBothWays = true;
    if (const char* e = getenv("HQ_CUT_BOTH_WAYS")) cutBothWays = atoi(e) != 0;
    cutVariants = 32;   // upper bound; the number tried per cut shrinks with the length of the gate list (trialsFor)
    if (const char* e = getenv("HQ_CUT_VARIANTS")) cutVariants = std::max(1, atoi(e));
    maxGroupGates = 384;
    if (const char* e = getenv("HQ_MAX_GROUP_GATES")) maxGroupGates = std::max(1, atoi(e));
}

// ---- stage split --------------------------------------------------------------------------------------
// The split is greedy, and a greedy split is fragile (one merged gate pair can turn two stages into three, i.e. one more exchange
// of the whole state): it is run with a few different tie-breaking orders and the result with the fewest stages -- then the
// fewest swapped qubits -- is kept.
std::vector<Compiler::Stage> Compiler::splitStages() const {
    std::vector<Stage> best;
    int bestSwapped = 1 << 30;
    const int nvar = MyGlobalVars::bit == 0 ? 1 : 3;
    std::vector<std::vector<Stage>> tried(nvar);
    for (int variant = 0; variant < nvar; variant++) tried[variant] = splitStagesVariant(variant);
    for (int variant = 0; variant < nvar; variant++) {
        std::vector<Stage>& st = tried[variant];
        int swapped = 0;
        for (size_t s = 1; s < st.size(); s++) swapped += bitCount(st[s].locals & ~st[s - 1].locals);
        if (best.empty() || st.size() < best.size() || (st.size() == best.size() && swapped < bestSwapped)) { best = std::move(st); bestSwapped = swapped; }
    }
    return best;
}

std::vector<Compiler::Stage> Compiler::splitStagesVariant(int variant) const {
    std::vector<Stage> stages;
    if (MyGlobalVars::bit == 0) {
        stages.push_back({gates, (qindex(1) << numQubits) - 1});
        return stages;
    }
    std::vector<int> remaining(gates.size());
    for (size_t i = 0; i < gates.size(); i++) remaining[i] = (int)i;
    qindex prevLocals = (qindex(1) << numLocal) - 1;
    const hyquas::FrontierScan scan(gates);
    std::vector<qindex> extra;
    std::vector<int> cand, count;
    while (!remaining.empty()) {
        // grow the local set greedily: repeatedly add the qubit that unlocks the most gates
        qindex locals = 0;
        while (bitCount(locals) < numLocal) {
            extra.assign(1, 0);   // candidate 0: nothing added
            cand.assign(1, -1);
            for (int qi = 0; qi < numQubits; qi++) {
                // variants differ in the order qubits are tried (ties go to the first one found)
                const int q = variant == 0 ? qi : (variant == 1 ? numQubits - 1 - qi : (qi + numQubits / 2) % numQubits);
                if (locals >> q & 1) continue;
                extra.push_back(qindex(1) << q);
                cand.push_back(q);
            }
            scan.counts(remaining, locals, 4096, extra, count);
            const int base = count[0];
            int best = -1, bestGain = base;
            for (size_t c = 1; c < cand.size(); c++) {
                const int q = cand[c], gain = count[c];
                // ties prefer qubits that are already local (fewer bits to swap)
                if (gain > bestGain || (gain == bestGain && best >= 0 && gain > base && (prevLocals >> q & 1) && !(prevLocals >> best & 1))) {
                    best = q; bestGain = gain;
                }
            }
            if (best < 0) break;
            locals |= qindex(1) << best;
        }
        // pad with currently-local qubits first, then anything
        for (int pass = 0; pass < 2 && bitCount(locals) < numLocal; pass++)
            for (int q = 0; q < numQubits && bitCount(locals) < numLocal; q++)
                if (!(locals >> q & 1) && (pass == 1 || (prevLocals >> q & 1))) locals |= qindex(1) << q;
        std::vector<int> take = scan.runnable(remaining, locals, 1 << 30);
        assert(!take.empty());
        Stage st; st.locals = locals;
        for (int gi : take) st.gates.push_back(gates[gi]);
        std::vector<int> rest;
        std::set_difference(remaining.begin(), remaining.end(), take.begin(), take.end(), std::back_inserter(rest));
        remaining.swap(rest);
        stages.push_back(std::move(st));
        prevLocals = locals;
    }
    if (stages.empty()) stages.push_back({{}, (qindex(1) << numLocal) - 1});
    return stages;
}

// ---- dense (TransMM-class) candidate ---------------------------------------------------------------------
// One launch of the fused dense kernel = up to 8 matrices of <= 6 qubits each whose qubits, together with the three
// lowest physical bits, fit one 12-bit tile.  Blocks are grown greedily from the frontier: repeatedly merge the qubits
// of the gate that lets the block absorb the most additional gates.
GateGroup Compiler::denseCandidate(const std::vector<Gate>& stageGates, const std::vector<int>& remaining0, const State& state,
                                   int nLocal, qindex exclude) const {
    const int nEff = nLocal - bitCount(exclude);   // qubits that vary in this launch
    Evaluator* ev = Evaluator::getInstance();
    GateGroup gg;
    gg.backend = Backend::BLAS;
    gg.state = state;
    std::vector<int> remaining = remaining0;
    const int lookahead = 512, tileCap = std::min(12, nEff);
    qindex tileQubits = 0;   // logical qubits whose physical bits the launch's tile must contain
    for (int p = 0; p < std::min(3, nLocal); p++) tileQubits |= qindex(1) << state.layout[p];
    auto qubitsOf = [](const Gate& g) {
        qindex q = qindex(1) << g.targetQubit;
        if (g.controlQubit >= 0) q |= qindex(1) << g.controlQubit;
        if (g.controlQubit2 >= 0) q |= qindex(1) << g.controlQubit2;
        return q;
    };
    double denseMs = 0;
    while ((int)gg.blocks.size() < 8 && !remaining.empty()) {
        // best block for every size cap; keep the one with the most gates per predicted millisecond.  The same qubit
        // sets come up again and again across caps and growth steps: their gate counts are memoised.
        qindex bestSet = 0; std::vector<int> bestTake; double bestRate = 0; int bestM = 0;
        std::unordered_map<qindex, int> memo;
        auto countFor = [&](qindex S2) {
            auto it = memo.find(S2);
            if (it != memo.end()) return it->second;
            int cnt = -1;
            if (bitCount(tileQubits | S2) <= tileCap) {
                bool local = true;
                for (int q = 0; q < numQubits; q++)
                    if ((S2 >> q & 1) && (state.pos[q] >= nLocal || (exclude >> state.pos[q] & 1))) local = false;
                if (local) cnt = (int)hyquas::runnableDense(stageGates, remaining, S2, lookahead).size();
            }
            memo.emplace(S2, cnt);
            return cnt;
        };
        std::vector<qindex> frontier;   // distinct qubit sets of the first gates
        {
            int scanned = 0;
            for (int gi : remaining) {
                if (++scanned > 96) break;
                const qindex q = qubitsOf(stageGates[gi]);
                if (std::find(frontier.begin(), frontier.end(), q) == frontier.end()) frontier.push_back(q);
            }
        }
        for (int cap = 3; cap <= matLimit; cap++) {
            qindex S = 0;
            int curCount = 0;
            while (true) {
                qindex pickSet = 0; int pickCount = curCount;
                for (qindex fq : frontier) {
                    const qindex S2 = S | fq;
                    if (S2 == S || bitCount(S2) > cap) continue;
                    const int cnt = countFor(S2);
                    if (cnt > pickCount || (cnt == pickCount && pickSet && cnt > curCount && bitCount(S2) < bitCount(pickSet))) {
                        pickSet = S2; pickCount = cnt;
                    }
                }
                if (!pickSet) break;
                S = pickSet;
                curCount = pickCount;
            }
            if (curCount <= 0) continue;
            const int m = std::max(3, bitCount(S));
            const double rate = curCount / ev->denseMs30[m];
            if (rate > bestRate * 1.0001) { bestRate = rate; bestSet = S; bestM = m; }
        }
        if (bestSet) bestTake = hyquas::runnableDense(stageGates, remaining, bestSet, lookahead);
        if (bestTake.empty()) break;
        const double blockMs = ev->denseMs30[bestM];
        // further blocks must pay for themselves: stop once the launch is compute-bound and the new block is slower per
        // gate than what the launch already achieves
        if (!gg.blocks.empty()) {
            const double sweep = ev->sweepMs30();
            const double before = std::max(sweep, denseMs), after = std::max(sweep, denseMs + blockMs);
            if ((after - before) / bestTake.size() > before / gg.gates.size()) break;
        }
        DenseBlock blk;
        blk.qubits = bestSet;
        for (int gi : bestTake) { blk.gates.push_back(stageGates[gi]); gg.gates.push_back(stageGates[gi]); }
        gg.blocks.push_back(std::move(blk));
        gg.relatedQubits |= bestSet;
        gg.matQubit = std::max(gg.matQubit, bestM);
        tileQubits |= bestSet;
        denseMs += blockMs;
        std::vector<int> rest;
        std::set_difference(remaining.begin(), remaining.end(), bestTake.begin(), bestTake.end(), std::back_inserter(rest));
        remaining.swap(rest);
    }
    std::vector<int> ms;
    for (auto& b : gg.blocks) ms.push_back(std::max(3, bitCount(b.qubits)));
    gg.predictedMs = gg.blocks.empty() ? 0 : ev->perfDense(nEff, ms);
    return gg;
}

// ---- gate groups inside one stage ----------------------------------------------------------------------
// `exclude`: local physical positions that do not vary in these launches (the swapped positions of a per-chunk group).
// The greedy cut is run over the stage's gates front to back AND back to front (commutation is symmetric, so a valid cut of
// the reversed list, reversed, is a valid cut of the list); the cheaper predicted total wins.  The front-to-back cut leaves
// its crumbs (launches with a handful of gates, a full sweep each) at the end of the stage, the other one at the start, and
// which of the two is smaller depends on the circuit.
// A tile-kernel launch costs max(sweep, arithmetic + exchanges): the greedy cut fills every group to the brim, which leaves some
// launches compute-bound (11 ms) next to sweep-bound ones with arithmetic to spare (5.5 ms).  Adjacent tile groups may trade
// gates without changing what is computed: a suffix-closed set of group i whose non-diagonal targets lie in group i+1's tile can
// open group i+1 instead, and a prefix-closed set of group i+1 that fits group i's tile can close group i.  Greedy descent on the
// predicted total, pair by pair; a group that ends up empty disappears (that is how trailing crumbs are absorbed).
void Compiler::rebalance(std::vector<GateGroup>& groups, int nEff) const {
    if (!rebalanceGroups || groups.size() < 2) return;
    Evaluator* ev = Evaluator::getInstance();
    auto cost = [&](const std::vector<Gate>& g) { return g.empty() ? 0.0 : ev->perfPerGate(nEff, g); };
    for (int pass = 0; pass < 4; pass++) {
        bool changed = false;
        for (size_t i = 0; i + 1 < groups.size(); i++) {
            GateGroup &A = groups[i], &B = groups[i + 1];
            if (A.backend != Backend::PerGate || B.backend != Backend::PerGate) continue;
            const double base = cost(A.gates) + cost(B.gates);
            double best = base - 0.05;   // a move must win at least 50 us
            int bestDir = 0;
            size_t bestK = 0;
            // direction +1: the tail of A opens B
            std::vector<int> order(A.gates.size());
            for (size_t j = 0; j < order.size(); j++) order[j] = (int)A.gates.size() - 1 - (int)j;
            const std::vector<int> tail = hyquas::runnableGates(A.gates, order, B.relatedQubits, 1 << 30);   // nearest the end first
            // direction -1: the head of B closes A
            std::vector<int> fwdOrder(B.gates.size());
            for (size_t j = 0; j < fwdOrder.size(); j++) fwdOrder[j] = (int)j;
            const std::vector<int> head = hyquas::runnableGates(B.gates, fwdOrder, A.relatedQubits, 1 << 30);
            auto split = [&](const std::vector<Gate>& src, const std::vector<int>& picked, size_t k, std::vector<Gate>& keep, std::vector<Gate>& moved) {
                std::vector<char> take(src.size(), 0);
                for (size_t j = 0; j < k; j++) take[picked[j]] = 1;
                keep.clear(); moved.clear();
                for (size_t j = 0; j < src.size(); j++) (take[j] ? moved : keep).push_back(src[j]);
            };
            std::vector<Gate> keep, moved, other;
            for (int dir : {+1, -1}) {
                const std::vector<int>& picked = dir > 0 ? tail : head;
                const std::vector<Gate>& src = dir > 0 ? A.gates : B.gates;
                const std::vector<Gate>& dst = dir > 0 ? B.gates : A.gates;
                // try a handful of sizes (every 4th gate, plus everything: emptying a group removes a whole sweep)
                for (size_t k = 1; k <= picked.size(); k = (k + 4 <= picked.size() || k == picked.size()) ? k + 4 : picked.size()) {
                    if ((int)(dst.size() + k) > maxGroupGates) break;
                    split(src, picked, k, keep, moved);
                    if (dir > 0) { other = moved; other.insert(other.end(), dst.begin(), dst.end()); }
                    else { other = dst; other.insert(other.end(), moved.begin(), moved.end()); }
                    const double c = cost(keep) + cost(other);
                    if (c < best) { best = c; bestDir = dir; bestK = k; }
                    if (k == picked.size()) break;
                }
            }
            if (!bestDir) continue;
            const std::vector<int>& picked = bestDir > 0 ? tail : head;
            if (bestDir > 0) {
                split(A.gates, picked, bestK, keep, moved);
                A.gates = keep;
                moved.insert(moved.end(), B.gates.begin(), B.gates.end());
                B.gates = moved;
            } else {
                split(B.gates, picked, bestK, keep, moved);
                B.gates = keep;
                A.gates.insert(A.gates.end(), moved.begin(), moved.end());
            }
            A.predictedMs = cost(A.gates);
            B.predictedMs = cost(B.gates);
            changed = true;
        }
        for (size_t i = groups.size(); i-- > 0;)
            if (groups[i].backend == Backend::PerGate && groups[i].gates.empty()) groups.erase(groups.begin() + i);
        if (!changed) break;
    }
}

// A launch with a handful of gates still costs a full sweep.  Its gates may be able to run EARLIER: gate g of a small group j can
// close group i < j when g's non-diagonal target lies in group i's tile and g commutes with everything that runs in between
// (the groups i+1 .. j-1 and the gates of group j that stay ahead of it).  If every gate of the small group finds such a home
// the group disappears and the schedule is one sweep shorter (qaoa_30: 7 -> 5 sweeps, supremacy_30: 10 -> 9).
static bool gatesCommute(const Gate& a, const Gate& b) {
    auto acts = [](const Gate& g, qindex& nd, qindex& d) {
        nd = d = 0;
        (g.isDiagonal() ? d : nd) |= qindex(1) << g.targetQubit;
        if (g.controlQubit >= 0) d |= qindex(1) << g.controlQubit;
        if (g.controlQubit2 >= 0) d |= qindex(1) << g.controlQubit2;
    };
    qindex an, ad, bn, bd;
    acts(a, an, ad);
    acts(b, bn, bd);
    return !(an & (bn | bd)) && !(bn & ad);
}

void Compiler::absorbCrumbs(std::vector<GateGroup>& groups, int nEff) const {
    if (!rebalanceGroups) return;
    Evaluator* ev = Evaluator::getInstance();
    const double sweep = ev->perfPerGate(nEff, std::vector<Gate>());
    for (size_t j = groups.size(); j-- > 1;) {
        if (groups[j].gates.size() > 24) continue;   // (a small dense launch is a crumb too; its gates can only move into tile groups)
        std::vector<std::vector<Gate>> added(j);          // gates appended to group i < j so far
        bool all = true;
        double extra = 0;
        for (const Gate& g : groups[j].gates) {
            int home = -1;
            for (int i = (int)j - 1; i >= 0; i--) {        // nearest group first
                // g would run at the end of group i: it must pass everything already placed after group i ...
                bool pass = true;
                for (size_t m = (size_t)i + 1; m < j && pass; m++) {
                    for (const Gate& o : groups[m].gates) if (!gatesCommute(g, o)) { pass = false; break; }
                    for (const Gate& o : added[m]) if (pass && !gatesCommute(g, o)) { pass = false; break; }
                }
                if (!pass) break;                          // ... and nothing further back can be reached either
                const bool fits = groups[i].backend == Backend::PerGate && (g.isDiagonal() || (groups[i].relatedQubits >> g.targetQubit & 1)) &&
                                  (int)(groups[i].gates.size() + added[i].size()) < maxGroupGates;
                if (fits) { home = i; break; }
                // not in this tile: g may still hop over group i as a whole if it commutes with all of it
                for (const Gate& o : groups[i].gates) if (!gatesCommute(g, o)) { pass = false; break; }
                for (const Gate& o : added[i]) if (pass && !gatesCommute(g, o)) { pass = false; break; }
                if (!pass) break;
            }
            if (home < 0) { all = false; break; }
            added[home].push_back(g);
        }
        if (!all) continue;
        for (size_t i = 0; i < j; i++) {
            if (added[i].empty()) continue;
            std::vector<Gate> merged = groups[i].gates;
            merged.insert(merged.end(), added[i].begin(), added[i].end());
            extra += ev->perfPerGate(nEff, merged) - groups[i].predictedMs;
        }
        if (extra >= std::max(groups[j].predictedMs, sweep)) continue;   // dearer than the sweep it saves
        for (size_t i = 0; i < j; i++) {
            if (added[i].empty()) continue;
            groups[i].gates.insert(groups[i].gates.end(), added[i].begin(), added[i].end());
            groups[i].predictedMs = ev->perfPerGate(nEff, groups[i].gates);
        }
        groups.erase(groups.begin() + j);
    }
}

// ---- search over greedy cuts, and its memory -------------------------------------------------------------
// The greedy cut is myopic: which of several equally good qubits joins a tile decides, many groups later, whether the stage needs
// one sweep more or less (supremacy_30: 10 launches / 69 ms predicted from the plain greedy, 9 / 63 ms from 2 of 40 differently
// seeded tie-breaks).  So a stage is cut with several seeds -- tile kernel only, priced by the evaluator -- and the three cheapest
// are finished (rebalance, crumbs) and compared with the plain cut; a schedule that is not predicted at least 1 % cheaper never
// replaces it.  Every trial costs about as much as the plain cut (1-3 ms), so which seed won is remembered: in the process, and in
// a small file next to the compiled kernels (one per circuit and partitioner configuration; the kernels of a new circuit take
// 0.3-0.6 s to compile, the search 20-80 ms).  A remembered seed is replayed, not trusted blindly: the cut it produces is valid by
// construction whatever the file says, and every rank derives the same one from the same bytes or from the same search.
int Compiler::trialsFor(size_t numGates) const {
    if (cutVariants <= 1 || numGates < 24) return 1;
    return (int)std::max<size_t>(1, std::min<size_t>((size_t)cutVariants, 16000 / numGates));
}

static unsigned long long fnv64(const void* data, size_t n, unsigned long long h) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 0x100000001b3ull; }
    return h;
}

unsigned long long Compiler::cutKey(const std::vector<Gate>& stageGates, const State& state, int nLocal, qindex exclude) const {
    unsigned long long h = 0xcbf29ce484222325ull;
    for (const Gate& g : stageGates) h = fnv64(&g.gateID, sizeof(int), h);   // ids are positions in this circuit's gate list
    h = fnv64(state.layout.data(), sizeof(int) * (size_t)nLocal, h);
    h = fnv64(&nLocal, sizeof(int), h);
    h = fnv64(&exclude, sizeof(exclude), h);
    return h;
}

// Everything the cuts of this circuit depend on: the gates, the machine shape, the partitioner's knobs and the evaluator's constants.
unsigned long long Compiler::circuitKey() const {
    unsigned long long h = fnv64("hqcut1", 6, 0xcbf29ce484222325ull);
    const int shape[] = {numQubits, numLocal, MyGlobalVars::numGPUs, MyGlobalVars::swapAnyBit ? 1 : 0, tileBits, pinnedBits, maxGroupGates,
                         rebalanceGroups ? 1 : 0, cutVariants, cutBothWays ? 1 : 0, enableOverlap ? 1 : 0, backendMode, matLimit};
    h = fnv64(shape, sizeof(shape), h);
    h = fnv64(&overlapSlack, sizeof(overlapSlack), h);
    const unsigned long long ev = Evaluator::getInstance()->signature(numLocal);
    h = fnv64(&ev, sizeof(ev), h);
    for (const Gate& g : gates) {
        const int head[4] = {(int)g.type, g.targetQubit, g.controlQubit, g.controlQubit2};
        h = fnv64(head, sizeof(head), h);
        h = fnv64(g.mat, sizeof(g.mat), h);
    }
    return h;
}

// (process-wide copy of what the files hold: a second compile of the same circuit in one process never touches the disk)
static std::mutex wisdomMutex;
static std::map<unsigned long long, std::unordered_map<unsigned long long, int>> wisdomOfCircuit;

static std::string wisdomPath(unsigned long long key) {
    char dir[4096];
    // (host-only runs -- the CPU tests -- keep their hands off the user's cache directory unless one is named explicitly)
    if ((MyGlobalVars::hostOnly && !getenv("HQ_JIT_CACHE")) || hq_cache_dir(dir, sizeof(dir)) != HQ_OK || !dir[0]) return std::string();
    char name[64];
    snprintf(name, sizeof(name), "/%016llx.cuts", key);
    return std::string(dir) + name;
}

void Compiler::loadWisdom() {
    wisdom.clear();
    wisdomDirty = false;
    if (cutVariants <= 1) return;
    const unsigned long long key = circuitKey();
    {
        std::lock_guard<std::mutex> lock(wisdomMutex);
        auto it = wisdomOfCircuit.find(key);
        if (it != wisdomOfCircuit.end()) { wisdom = it->second; return; }
    }
    const std::string path = wisdomPath(key);
    if (path.empty()) return;
    if (FILE* f = fopen(path.c_str(), "r")) {
        unsigned long long k;
        int seed;
        while (fscanf(f, "%llx %d", &k, &seed) == 2)
            if (seed >= 0 && seed < cutVariants) wisdom[k] = seed;   // (anything else in the file is ignored: the cut is searched again)
        fclose(f);
    }
}

void Compiler::saveWisdom() const {
    if (!wisdomDirty || cutVariants <= 1) return;
    wisdomDirty = false;
    const unsigned long long key = circuitKey();
    {
        std::lock_guard<std::mutex> lock(wisdomMutex);
        if (wisdomOfCircuit.size() > 256) wisdomOfCircuit.clear();   // (a long-lived process compiling ever new circuits)
        wisdomOfCircuit[key] = wisdom;
    }
    const std::string path = wisdomPath(key);
    if (path.empty()) return;
    const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
    FILE* f = fopen(tmp.c_str(), "w");
    if (!f) return;
    for (const auto& kv : wisdom) fprintf(f, "%016llx %d\n", kv.first, kv.second);
    const bool ok = fclose(f) == 0;
    if (!ok || rename(tmp.c_str(), path.c_str()) != 0) unlink(tmp.c_str());
}

std::vector<GateGroup> Compiler::cutGroups(const std::vector<Gate>& stageGates, const State& state, int nLocal, qindex exclude,
                                           bool search) const {
    const int nEff = nLocal - bitCount(exclude);
    auto finish = [&](std::vector<GateGroup>& out) {
        rebalance(out, nEff);
        const size_t before = out.size();
        absorbCrumbs(out, nEff);
        if (out.size() != before) rebalance(out, nEff);
    };
    auto total = [](const std::vector<GateGroup>& gs) { double t = 0; for (auto& g : gs) t += g.predictedMs; return t; };
    const int V = search ? trialsFor(stageGates.size()) : 1;
    const unsigned long long key = V > 1 ? cutKey(stageGates, state, nLocal, exclude) : 0;
    int known = -1;
    if (V > 1) {
        auto it = wisdom.find(key);
        if (it != wisdom.end()) known = it->second;
    }
    if (known > 0) {   // the seed that won this very cut before
        std::vector<GateGroup> out = cutGroupsGreedy(stageGates, state, nLocal, exclude, known, true);
        finish(out);
        return out;
    }
    std::vector<GateGroup> best = cutGroupsBothWays(stageGates, state, nLocal, exclude);
    finish(best);
    if (V <= 1 || known == 0) return best;
    int winner = 0;
    bool tileOnly = best.size() >= 3;   // (one or two launches: nothing to gain)
    for (auto& g : best) if (g.backend != Backend::PerGate) tileOnly = false;   // dense-friendly circuits keep the hybrid cut
    if (tileOnly) {
        struct Trial { double pre; int seed; std::vector<GateGroup> groups; };
        std::vector<Trial> trials;
        for (int v = 1; v < V; v++) {
            Trial t;
            t.seed = v;
            t.groups = cutGroupsGreedy(stageGates, state, nLocal, exclude, v, true);
            t.pre = total(t.groups);
            trials.push_back(std::move(t));
        }
        std::stable_sort(trials.begin(), trials.end(), [](const Trial& a, const Trial& b) { return a.pre < b.pre; });
        double bestMs = total(best);
        const double mustBeat = bestMs * 0.99;
        for (size_t i = 0; i < trials.size() && i < 3; i++) {
            finish(trials[i].groups);
            const double ms = total(trials[i].groups);
            if (ms < mustBeat && ms < bestMs) { bestMs = ms; winner = trials[i].seed; best.swap(trials[i].groups); }
        }
    }
    wisdom[key] = winner;
    wisdomDirty = true;
    return best;
}

std::vector<GateGroup> Compiler::cutGroupsBothWays(const std::vector<Gate>& stageGates, const State& state, int nLocal, qindex exclude) const {
    std::vector<GateGroup> fwd = cutGroupsGreedy(stageGates, state, nLocal, exclude);
    // (only for stages of <= 256 gates: there the second cut costs about a millisecond of compile time and removes a sweep
    // from bv / adder; on the long random circuits it never won and would only add 15-60 ms to compile())
    if (!cutBothWays || stageGates.size() < 2 || stageGates.size() > 256 || fwd.size() < 2) return fwd;
    std::vector<Gate> rev(stageGates.rbegin(), stageGates.rend());
    std::vector<GateGroup> bwd = cutGroupsGreedy(rev, state, nLocal, exclude);
    auto total = [](const std::vector<GateGroup>& gs) { double t = 0; for (auto& g : gs) t += g.predictedMs; return t; };
    // fewer sweeps, or clearly cheaper: a 2-3 % predicted edge is inside the evaluator's error and cost hidden_shift_36 its
    // overlap group (611 vs 564 ms measured on 8 GPUs, profiles r01_s25 vs r01_s18)
    if (!(bwd.size() < fwd.size() || total(bwd) < total(fwd) * 0.95)) return fwd;
    std::reverse(bwd.begin(), bwd.end());
    for (GateGroup& g : bwd) {
        std::reverse(g.gates.begin(), g.gates.end());
        std::reverse(g.blocks.begin(), g.blocks.end());
        for (DenseBlock& b : g.blocks) std::reverse(b.gates.begin(), b.gates.end());
    }
    return bwd;
}

std::vector<GateGroup> Compiler::cutGroupsGreedy(const std::vector<Gate>& stageGates, const State& state, int nLocal, qindex exclude,
                                                 int variant, bool tileOnly) const {
    const int nEff = nLocal - bitCount(exclude);
    std::vector<GateGroup> groups;
    std::vector<int> remaining(stageGates.size());
    for (size_t i = 0; i < stageGates.size(); i++) remaining[i] = (int)i;
    const int K = std::min(tileBits, nEff), C = std::min(pinnedBits, K);
    const int lookahead = 2048;
    const hyquas::FrontierScan scan(stageGates);
    static const bool checkScans = getenv("HQ_CHECK_SCANS") != nullptr;
    // variant 0 always takes the qubit with the largest gain (first one found on ties); the others take a random one among the
    // qubits within 0-2 gates of the largest gain, from a generator seeded with the variant number (every rank draws the same)
    unsigned long long rng = 0x9E3779B97F4A7C15ull * (unsigned long long)(variant + 1);
    auto draw = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
    std::vector<qindex> extra;
    std::vector<int> cand, count;
    while (!remaining.empty()) {
        qindex tile = 0;
        for (int p = 0; p < C; p++) tile |= qindex(1) << state.layout[p];
        while (bitCount(tile) < K) {
            extra.assign(1, 0);
            cand.assign(1, -1);
            for (int p = C; p < nLocal; p++) {
                const int q = state.layout[p];
                if ((tile >> q & 1) || (exclude >> p & 1)) continue;
                extra.push_back(qindex(1) << q);
                cand.push_back(q);
            }
            scan.counts(remaining, tile, lookahead, extra, count);
            if (checkScans)   // (tests: the one-pass scan must agree with the reference scan, candidate by candidate)
                for (size_t c = 0; c < cand.size(); c++)
                    if ((size_t)count[c] != hyquas::runnableGates(stageGates, remaining, tile | extra[c], lookahead).size()) {
                        fprintf(stderr, "FrontierScan disagrees with runnableGates\n");
                        abort();
                    }
            const int cur = count[0];
            int best = -1, bestGain = cur;
            for (size_t c = 1; c < cand.size(); c++) if (count[c] > bestGain) { best = cand[c]; bestGain = count[c]; }
            if (best < 0) break;
            if (variant > 0) {
                const int slack = (int)(draw() % 3);
                int eligible = 0;
                for (size_t c = 1; c < cand.size(); c++) if (count[c] > cur && count[c] + slack >= bestGain) eligible++;
                int pick = (int)(draw() % (unsigned long long)eligible);
                for (size_t c = 1; c < cand.size(); c++)
                    if (count[c] > cur && count[c] + slack >= bestGain && pick-- == 0) { best = cand[c]; break; }
            }
            tile |= qindex(1) << best;
        }
        for (int p = 0; p < nLocal && bitCount(tile) < K; p++)   // pad with the lowest free physical bits (longer runs)
            if (!(tile >> state.layout[p] & 1) && !(exclude >> p & 1)) tile |= qindex(1) << state.layout[p];
        std::vector<int> take = scan.runnable(remaining, tile, lookahead);
        if ((int)take.size() > maxGroupGates) take.resize(maxGroupGates);   // a prefix of a runnable set is runnable
        assert(!take.empty());
        GateGroup gg;
        gg.backend = Backend::PerGate;
        gg.relatedQubits = tile;
        gg.state = state;
        for (int gi : take) gg.gates.push_back(stageGates[gi]);
        gg.predictedMs = Evaluator::getInstance()->perfPerGate(nEff, gg.gates);
        // A tile launch that is (nearly) sweep-bound cannot lose to a dense launch: that one costs at least a sweep too and, needing
        // every operand of its gates inside its blocks, never holds more gates.  Growing the dense candidate is the expensive part
        // of the hybrid cut (6 of 10 ms for supremacy_30), so it is only done for compute-bound tile groups.
        const bool tileIsSweepBound = gg.predictedMs <= 1.35 * Evaluator::getInstance()->perfPerGate(nEff, std::vector<Gate>());
        if (!tileOnly && backendMode != 1 && nEff >= 8 && (backendMode == 3 || nEff < 10 || !tileIsSweepBound)) {
            // hybrid choice (the reference's AdvanceCompiler::run, src/compiler.cpp:250-278): price a dense launch for
            // the same frontier and keep whichever costs fewer predicted milliseconds per gate
            GateGroup dense = denseCandidate(stageGates, remaining, state, nLocal, exclude);
            const bool pick = !dense.gates.empty() &&
                (backendMode == 3 || nEff < 10 ||
                 dense.predictedMs / dense.gates.size() < gg.predictedMs / gg.gates.size());
            if (pick) {
                gg = std::move(dense);
                take.clear();
                std::vector<int> ids;
                for (auto& g : gg.gates) ids.push_back(g.gateID);
                std::sort(ids.begin(), ids.end());
                for (int gi : remaining) if (std::binary_search(ids.begin(), ids.end(), stageGates[gi].gateID)) take.push_back(gi);
            }
        }
        std::vector<int> rest;
        std::set_difference(remaining.begin(), remaining.end(), take.begin(), take.end(), std::back_inserter(rest));
        remaining.swap(rest);
        groups.push_back(std::move(gg));
    }
    return groups;
}

Schedule Compiler::run() {
    Schedule schedule;
    State state(numQubits);
    loadWisdom();
    std::vector<Stage> stages = splitStages();
    // pass 1: layouts and swaps (independent of how stages are cut into groups)
    for (size_t s = 0; s < stages.size(); s++) {
        LocalGroup lg;
        lg.relatedQubits = stages[s].locals;
        if (s == 0) {
            // |0...0> is invariant under qubit relabelling: just declare the stage-0 locals to be at [0, numLocal)
            // Stage-0 locals that leave later go to the TOP local positions, earliest leavers highest: the qubits coming in
            // land on those same positions, so the traffic between local and global keeps using high positions (long
            // contiguous runs on the wire, and for the nccl transport no local bit swaps at the first exchange).
            State st(numQubits);
            std::vector<int> stay, leave;
            std::vector<int> firstOut(numQubits, 1 << 30);
            for (int q = 0; q < numQubits; q++)
                for (size_t t = 1; t < stages.size(); t++)
                    if ((stages[0].locals >> q & 1) && !(stages[t].locals >> q & 1)) { firstOut[q] = (int)t; break; }
            for (int q = 0; q < numQubits; q++)
                if (stages[0].locals >> q & 1) (firstOut[q] < (1 << 30) ? leave : stay).push_back(q);
            std::stable_sort(leave.begin(), leave.end(), [&](int a, int b) { return firstOut[a] > firstOut[b]; });
            int p = 0;
            for (int q : stay) { st.pos[q] = p; st.layout[p] = q; p++; }
            for (int q : leave) { st.pos[q] = p; st.layout[p] = q; p++; }
            for (int q = 0; q < numQubits; q++)
                if (!(stages[0].locals >> q & 1)) { st.pos[q] = p; st.layout[p] = q; p++; }
            state = st;
        } else {
            lg.swap = hyquas::planSwap(state, stages[s].locals, numQubits, numLocal, MyGlobalVars::swapAnyBit);
        }
        lg.state = state;
        schedule.localGroups.push_back(std::move(lg));
    }
    // pass 2: work to hide the exchange behind.  The stage split is greedy, so the head of stage s always needs the
    // incoming qubits; what CAN run while chunks are still on the wire is the tail of stage s-1 (the reference's moveToNext
    // + overlapGroups, src/compiler.cpp:34-68, src/executor.cpp:41-47), run chunk by chunk as the chunks land.
    std::map<size_t, std::vector<GateGroup>> readyCuts;   // stage -> its full groups, when pass 2 already cut exactly that gate list
    for (size_t s = 1; s < stages.size() && enableOverlap; s++) {
        LocalGroup& lg = schedule.localGroups[s];
        const int k = (int)lg.swap.localBit.size();
        if (k == 0 || numLocal - k < 10) continue;   // a per-chunk launch needs a 10-bit tile of varying positions
        qindex lowSet = 0, exclude = 0;
        for (int b : lg.swap.localBit) exclude |= qindex(1) << b;
        for (int p = 0; p < numLocal; p++) if (!(exclude >> p & 1)) lowSet |= qindex(1) << lg.state.layout[p];
        std::vector<Gate>& prev = stages[s - 1].gates;
        Evaluator* ev = Evaluator::getInstance();
        const double commMs = ev->perfSwap(numLocal, k);
        auto total = [](const std::vector<GateGroup>& gs) { double t = 0; for (auto& g : gs) t += g.predictedMs; return t; };
        // Gate-granular deferral (the reference's moveToNext, src/compiler.cpp:34-68): the maximal suffix-closed set of stage
        // s-1's gates whose non-diagonal targets stay local and off the swapped positions, re-cut into per-chunk groups in the
        // NEW layout.  Any suffix of that group sequence is itself suffix-closed, so the choice is how many trailing groups
        // to defer: the one that minimises the predicted  (what is left of stage s-1) + max(exchange, deferred work),
        // i.e. deferral is taken only where it does not cost more in extra sweeps than it hides.
        std::vector<int> order(prev.size());
        for (size_t i = 0; i < order.size(); i++) order[i] = (int)prev.size() - 1 - (int)i;   // scan from the end
        std::vector<int> tail = hyquas::runnableGates(prev, order, lowSet, 1 << 30);
        std::sort(tail.begin(), tail.end());
        std::vector<Gate> tailGates;
        for (int gi : tail) tailGates.push_back(prev[gi]);
        if (tailGates.empty()) continue;
        // (the plain cut here: which gates CAN be deferred should not depend on seeds, so that every option below is the plain
        // pipeline's option with its full groups cut at least as well)
        std::vector<GateGroup> cand = cutGroups(tailGates, lg.state, numLocal, exclude, false);
        const State& prevState = schedule.localGroups[s - 1].state;
        // Per-chunk launches leave 32 of 148 SMs and some HBM bandwidth to the exchange kernel.  Measured (r02_m8, supremacy_33:
        // 1/8-state launches of 6.3 ms groups take 1.25 ms, not 0.79) the slowdown is ~1.5; the chooser was measured with 1.3
        // (8 GPUs: hidden_frac 0.42-0.65) and with 1.55 (4 GPUs: 0.24-0.47): the more eager setting hides more, it stays.
        const double underExchange = 1.3;
        // Chunks land one after the other (the one that stays at once, then one per exchange step); the deferred work of a chunk
        // starts when the chunk has landed and the previous chunk's work is done, so the last chunk's share is always exposed:
        // with k = 1 half of the deferred work cannot be hidden, with k = 3 an eighth (r02_m2: supremacy_31 on 2 GPUs lost
        // 4 ms to a deferral that a max(exchange, work) model had priced as a 5 ms win).
        auto pipelined = [&](double workMs) {
            const int chunks = 1 << k;
            const double w = workMs / chunks;
            double done = 0;
            for (int c = 0; c < chunks; c++) {
                const double landed = commMs * c / (chunks - 1);
                done = std::max(done, landed) + w;
            }
            return done;
        };
        // every candidate costs one cut of stage s-1 (milliseconds of compile time): the cut that wins is kept for pass 3, and at
        // most four deferral depths are tried
        static const bool verbosePlan = getenv("HQ_PLAN_VERBOSE") != nullptr;   // developer aid: the options weighed below
        std::vector<GateGroup> bestCut = cutGroups(prev, prevState, numLocal, 0);
        double bestCost = total(bestCut) + commMs;
        if (verbosePlan) fprintf(stderr, "[hq plan] stage %zu: no deferral: %zu groups %.2f ms + exchange %.2f ms\n", s, bestCut.size(), total(bestCut), commMs);
        size_t bestFirst = cand.size();
        double deferredMs = 0;
        for (size_t first = cand.size(); first-- > 0 && cand.size() - first <= 4;) {
            deferredMs += cand[first].predictedMs * (1 << k) * underExchange;   // predictedMs is per chunk
            if (deferredMs > commMs * overlapSlack * 2.0) break;
            std::vector<int> ids;
            for (size_t i = first; i < cand.size(); i++) for (auto& g : cand[i].gates) ids.push_back(g.gateID);
            std::sort(ids.begin(), ids.end());
            std::vector<Gate> rest;
            for (auto& g : prev) if (!std::binary_search(ids.begin(), ids.end(), g.gateID)) rest.push_back(g);
            std::vector<GateGroup> cut = cutGroups(rest, prevState, numLocal, 0);
            const double cost = total(cut) + pipelined(deferredMs);
            if (verbosePlan)
                fprintf(stderr, "[hq plan] stage %zu: deferring %zu group(s): rest %zu groups %.2f ms + pipelined(%.2f) = %.2f -> %.2f ms; best so far %.2f\n",
                        s, cand.size() - first, cut.size(), total(cut), deferredMs, pipelined(deferredMs), cost, bestCost);
            // (a huge HQ_OVERLAP_SLACK forces the maximal deferral whatever it costs: tests of the per-chunk path at sizes
            // whose exchange is too short to be worth hiding)
            if (cost < bestCost - 1e-9 || overlapSlack > 1e6) { bestCost = cost; bestFirst = first; bestCut.swap(cut); }
        }
        readyCuts[s - 1] = std::move(bestCut);   // stage s-1 is final now: pass 3 need not cut it again
        if (bestFirst == cand.size()) continue;
        std::vector<int> ids;
        for (size_t i = bestFirst; i < cand.size(); i++) {
            for (auto& g : cand[i].gates) ids.push_back(g.gateID);
            lg.overlapGroups.push_back(cand[i]);
        }
        std::sort(ids.begin(), ids.end());
        std::vector<Gate> rest;
        for (auto& g : prev) if (!std::binary_search(ids.begin(), ids.end(), g.gateID)) rest.push_back(g);
        prev.swap(rest);
    }
    // pass 3: cut what is left of every stage into full groups
    for (size_t s = 0; s < stages.size(); s++) {
        auto it = readyCuts.find(s);
        if (it != readyCuts.end()) schedule.localGroups[s].fullGroups = std::move(it->second);
        else schedule.localGroups[s].fullGroups = cutGroups(stages[s].gates, schedule.localGroups[s].state, numLocal, 0);
    }
    schedule.finalState = state;
    saveWisdom();
    return schedule;
}

namespace hyquas {
// Independent check of a schedule against the gate list it was cut from (tests; works without a GPU).  The launches, in the order
// the executor issues them (per stage: the per-chunk groups under the incoming exchange, then the full groups), must hold every
// gate exactly once, keep the relative order of every pair of gates that does not commute, and give every non-diagonal gate a
// target that is local in its stage, inside the launch's tile, and -- for per-chunk groups -- off the positions being exchanged.
std::string checkSchedule(const Schedule& schedule, const std::vector<Gate>& gates, int numQubits, int numLocal, int tileBits) {
    char msg[256];
    std::vector<int> where(gates.size(), -1);   // gate id -> position in the launch order
    int next = 0, stage = 0;
    for (const LocalGroup& lg : schedule.localGroups) {
        qindex swapped = 0;
        for (int b : lg.swap.localBit) swapped |= qindex(1) << b;
        int which = 0;
        for (const std::vector<GateGroup>* groups : {&lg.overlapGroups, &lg.fullGroups}) {
            const bool perChunk = which++ == 0;
            for (const GateGroup& gg : *groups) {
                if (gg.backend == Backend::PerGate && bitCount(gg.relatedQubits) > tileBits) return "a tile of more than tileBits qubits";
                qindex blockQubits = 0;
                for (const DenseBlock& b : gg.blocks) blockQubits |= b.qubits;
                for (const Gate& g : gg.gates) {
                    if (g.gateID < 0 || g.gateID >= (int)gates.size()) return "a gate id outside the circuit";
                    if (where[g.gateID] >= 0) { snprintf(msg, sizeof(msg), "gate %d is scheduled twice", g.gateID); return msg; }
                    where[g.gateID] = next++;
                    qindex need = g.isDiagonal() ? 0 : qindex(1) << g.targetQubit;   // qubits that must be movable in this launch
                    if (gg.backend == Backend::BLAS) {
                        need |= qindex(1) << g.targetQubit;
                        if (g.controlQubit >= 0) need |= qindex(1) << g.controlQubit;
                        if (g.controlQubit2 >= 0) need |= qindex(1) << g.controlQubit2;
                        if (need & ~blockQubits) { snprintf(msg, sizeof(msg), "gate %d reaches outside its dense blocks", g.gateID); return msg; }
                    } else if (need & ~gg.relatedQubits) {
                        snprintf(msg, sizeof(msg), "stage %d: target of gate %d is not in its launch's tile", stage, g.gateID);
                        return msg;
                    }
                    for (int q = 0; q < numQubits; q++) {
                        if (!(need >> q & 1)) continue;
                        const int p = lg.state.pos[q];
                        if (p >= numLocal) { snprintf(msg, sizeof(msg), "stage %d: gate %d needs qubit %d, which is global", stage, g.gateID, q); return msg; }
                        if (perChunk && (swapped >> p & 1)) { snprintf(msg, sizeof(msg), "stage %d: per-chunk gate %d needs a position under exchange", stage, g.gateID); return msg; }
                    }
                }
            }
        }
        stage++;
    }
    for (size_t i = 0; i < gates.size(); i++)
        if (where[i] < 0) { snprintf(msg, sizeof(msg), "gate %zu is not scheduled", i); return msg; }
    for (size_t j = 0; j < gates.size(); j++)
        for (size_t i = 0; i < j; i++)
            if (where[i] > where[j] && !gatesCommute(gates[i], gates[j])) {
                snprintf(msg, sizeof(msg), "gates %zu and %zu do not commute but run in the opposite order", i, j);
                return msg;
            }
    return std::string();
}
}  // namespace hyquas

