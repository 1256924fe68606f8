// Per-group specialisation of the gate-group kernel: plan tables -> straight-line CUDA C++ source.
//
// group_kernel.cu interprets a plan (rounds + lowered op list) with one indexed jump per gate; ncu (profiles/
// r01_s15) showed that decode chain, not FP64 or HBM, bounding gate-heavy launches (FP64 pipe 36 % active).  Here the same
// plan is turned into the source of ONE kernel that does exactly this group: the tile pipeline (persistent CTAs, TMA bulk
// tile loads, rounds with 16 amplitudes per thread in registers, swizzled shared-memory exchanges between rounds) is the
// same, but every gate is emitted as the FP64 instructions it needs and nothing else.  The emitter is a small symbolic
// compiler over the 32 scalars a thread holds:
//   * every scalar is (static coefficient) x (variable).  A product by a constant only changes the coefficient; a sum
//     c1*x + c2*y is emitted as ONE fma, x + (c2/c1)*y, and carries c1 on as its coefficient.  Hence: X / Y / Z / S / CNOT / CZ on
//     register qubits cost nothing (renaming), a butterfly (H, RX/RY(+-pi/2)) or any real 2x2 or any diagonal phase costs
//     one FP64 instruction per scalar, a general complex 2x2 three.  Coefficients are flushed (one multiply per scalar, less
//     a common factor that is deferred to the launch's scalar) when the round stores its amplitudes;
//   * controls on register qubits are resolved at emit time (only the affected register indices get code); controls and
//     diagonal targets on thread / outside bits become branches on `tid` / the tile base;
//   * diagonal runs (per-thread factor) are evaluated at run time per thread; runs of +-1 phases (CZ / Z) are applied as sign
//     flips on the integer pipe.
// The source is compiled by NVRTC for sm_100a (group_jit_rt.cpp) and cached by content hash in memory and on disk.
//
// The same emitter can produce a HOST flavour (plain C++, threads and tiles replayed serially) that the CPU tests compile with
// g++ and compare with the oracle: the arithmetic text of the two flavours is identical, only the skeleton differs.
//
// Reference role: doCompute's per-gate switch, src/kernelOpt.cu:214-386 (one SMEM read-modify-write + barrier per gate there).
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "group_jit.h"
#include "group_plan.h"
#include "hq_internal.h"

namespace hq {
namespace {

typedef std::complex<double> cplx;

std::string lit(double c) {
    char buf[64];
    if (c == 0.0) return "0.0";
    if (c == 1.0) return "1.0";
    if (c == -1.0) return "(-1.0)";
    snprintf(buf, sizeof(buf), c < 0 ? "(%a)" : "%a", c);
    return buf;
}
std::string hex64(uint64_t v) {
    char buf[32];
    snprintf(buf, sizeof(buf), "0x%llxull", (unsigned long long)v);
    return buf;
}
std::string hex32(uint32_t v) {
    char buf[32];
    snprintf(buf, sizeof(buf), "0x%xu", v);
    return buf;
}

// value = sum over bits b of ((src >> from[b]) & 1) << to[b], written as a few shift-and-mask terms (runs of consecutive
// bits that keep their distance are merged)
std::string deposit(const std::string& src, const std::vector<std::pair<int, int>>& from_to, bool wide) {
    std::string out;
    size_t i = 0;
    while (i < from_to.size()) {
        size_t j = i + 1;
        while (j < from_to.size() && from_to[j].first == from_to[j - 1].first + 1 && from_to[j].second == from_to[j - 1].second + 1) ++j;
        const int f = from_to[i].first, t = from_to[i].second, n = (int)(j - i);
        const uint64_t m = ((1ull << n) - 1) << f;
        std::string term = "(" + src + " & " + (wide ? hex64(m) : hex32((uint32_t)m)) + ")";
        if (wide) term = "(u64)" + term;
        if (t > f) term = "(" + term + " << " + std::to_string(t - f) + ")";
        else if (t < f) term = "(" + term + " >> " + std::to_string(f - t) + ")";
        out += (out.empty() ? "" : " | ") + term;
        i = j;
    }
    return out.empty() ? (wide ? "0ull" : "0u") : out;
}

struct Term { int var; double c; };

struct Emitter {
    const hq_group_plan& plan;
    const bool host;
    bool zero_in = false;   // the launch's input is |0...0> (amplitude 0 = 1 on the rank that holds it): nothing is read from HBM
    std::string o;          // the source being written
    int nvar = 0;
    Term re[R], im[R];
    int K, NT, TILE;
    int tile_phys[16];      // tile bit -> physical bit
    int phys_tile[64];      // physical bit -> tile bit or -1
    const DevRound* rounds;
    const DevOp* ops;
    // per round
    int thread_of_tile[16]; // tile bit -> thread-id bit or -1
    double deferred = 1.0;  // product of the common factors left out of earlier rounds' stores
    int stat_fp = 0;        // FP64 instructions emitted per thread (statistics)
    int stat_fused = 0;     // blocks emitted as one product matrix
    bool ok = true;         // false: the plan holds something the emitter does not understand (caller falls back)

    Emitter(const hq_group_plan& p, bool h) : plan(p), host(h) {
        K = plan.K; NT = plan.NT; TILE = 1 << K;
        for (int i = 0; i < 64; ++i) phys_tile[i] = -1;
        for (int b = 0, k = 0; b < plan.L; ++b)
            if (plan.tile_mask >> b & 1) { tile_phys[k] = b; phys_tile[b] = k++; }
        rounds = reinterpret_cast<const DevRound*>(plan.blob.data() + plan.o_rounds);
        ops = reinterpret_cast<const DevOp*>(plan.blob.data() + plan.o_ops);
    }

    void line(const std::string& s) { o += "        " + s + "\n"; }
    std::string V(int v) const { return "v" + std::to_string(v); }
    int fresh() { return nvar++; }

    // ---- symbolic arithmetic -------------------------------------------------------------------------------------
    static Term sc(Term t, double k) { return Term{t.var, t.c * k}; }

    // sum of terms -> one variable with a pending coefficient: c1*x1 + c2*x2 + ... = c1 * (x1 + (c2/c1) x2 + ...)
    Term lincomb(std::vector<Term> ts) {
        ts.erase(std::remove_if(ts.begin(), ts.end(), [](const Term& t) { return t.c == 0.0; }), ts.end());
        if (ts.empty()) {
            const int v = fresh();
            line("double " + V(v) + " = 0.0;");
            return Term{v, 1.0};
        }
        if (ts.size() == 1) return ts[0];
        const double g = ts[0].c;
        std::string expr = V(ts[0].var);
        for (size_t j = 1; j < ts.size(); ++j) {
            const double k = ts[j].c / g;
            if (k == 1.0) expr = "(" + expr + " + " + V(ts[j].var) + ")";
            else if (k == -1.0) expr = "(" + expr + " - " + V(ts[j].var) + ")";
            else expr = "fma(" + lit(k) + ", " + V(ts[j].var) + ", " + expr + ")";
            ++stat_fp;
        }
        const int v = fresh();
        line("double " + V(v) + " = " + expr + ";");
        return Term{v, g};
    }

    void mul_amp(int i, cplx d) {   // amplitude i *= d
        const Term x = re[i], y = im[i];
        re[i] = lincomb({sc(x, d.real()), sc(y, -d.imag())});
        im[i] = lincomb({sc(y, d.real()), sc(x, d.imag())});
    }

    // 2x2 complex matrix on register bit tb, restricted to the register indices that contain creg
    void apply2x2(const cplx M[4], int tb, uint32_t creg) {
        for (int p = 0; p < R / 2; ++p) {
            const int lo = ((p >> tb) << (tb + 1)) | (p & ((1 << tb) - 1)), hi = lo | (1 << tb);
            if ((lo & creg) != creg) continue;
            const Term lr = re[lo], li = im[lo], hr = re[hi], hi_ = im[hi];
            const Term nlr = lincomb({sc(lr, M[0].real()), sc(li, -M[0].imag()), sc(hr, M[1].real()), sc(hi_, -M[1].imag())});
            const Term nli = lincomb({sc(li, M[0].real()), sc(lr, M[0].imag()), sc(hi_, M[1].real()), sc(hr, M[1].imag())});
            const Term nhr = lincomb({sc(hr, M[3].real()), sc(hi_, -M[3].imag()), sc(lr, M[2].real()), sc(li, -M[2].imag())});
            const Term nhi = lincomb({sc(hi_, M[3].real()), sc(hr, M[3].imag()), sc(li, M[2].real()), sc(lr, M[2].imag())});
            re[lo] = nlr; im[lo] = nli; re[hi] = nhr; im[hi] = nhi;
        }
    }

    // ---- run-time predicates ---------------------------------------------------------------------------------------
    // all bits of `mask` (physical positions, none on a register qubit of this round) are 1
    std::string pred_all(uint64_t mask) const {
        uint64_t outside = 0; uint32_t tmask = 0;
        for (int b = 0; b < 64; ++b) {
            if (!(mask >> b & 1)) continue;
            if (phys_tile[b] < 0) outside |= 1ull << b;
            else if (thread_of_tile[phys_tile[b]] < 0) const_cast<Emitter*>(this)->ok = false;   // a register qubit: planner bug
            else tmask |= 1u << thread_of_tile[phys_tile[b]];
        }
        std::string s;
        if (outside) s = "((tbase & " + hex64(outside) + ") == " + hex64(outside) + ")";
        if (tmask) s += std::string(s.empty() ? "" : " && ") + "((tid & " + hex32(tmask) + ") == " + hex32(tmask) + ")";
        return s;   // empty: always true
    }
    static std::string both(const std::string& a, const std::string& b) {
        if (a.empty()) return b;
        if (b.empty()) return a;
        return a + " && " + b;
    }

    // ---- conditional sections: the static state after the section must equal the state before it -----------------------
    Term snap_re[R], snap_im[R];
    void cond_begin(const std::string& cond) {
        std::memcpy(snap_re, re, sizeof(re));
        std::memcpy(snap_im, im, sizeof(im));
        o += "        if (" + cond + ") {\n";
    }
    void cond_end() {
        std::vector<std::pair<int, std::string>> assign;   // (old variable, temporary holding its new value)
        auto fix = [&](const Term& now, const Term& was) {
            if (now.var == was.var && now.c == was.c) return;
            const double k = now.c / was.c;
            const int t = fresh();
            if (k == 1.0) line("double " + V(t) + " = " + V(now.var) + ";");
            else if (k == -1.0) line("double " + V(t) + " = -" + V(now.var) + ";");
            else { line("double " + V(t) + " = " + lit(k) + " * " + V(now.var) + ";"); ++stat_fp; }
            assign.push_back({was.var, V(t)});
        };
        for (int i = 0; i < R; ++i) { fix(re[i], snap_re[i]); fix(im[i], snap_im[i]); }
        for (auto& a : assign) line(V(a.first) + " = " + a.second + ";");
        o += "        }\n";
        std::memcpy(re, snap_re, sizeof(re));
        std::memcpy(im, snap_im, sizeof(im));
    }

    // ---- ops ---------------------------------------------------------------------------------------------------------
    static void decode2x2(const DevOp& op, uint32_t kind, cplx M[4]) {
        const double* m = op.m;
        switch (kind) {
            case OP_GEN: for (int i = 0; i < 4; ++i) M[i] = cplx(m[2 * i], m[2 * i + 1]); break;
            case OP_REAL: for (int i = 0; i < 4; ++i) M[i] = cplx(m[i], 0.0); break;
            case OP_RXL: M[0] = cplx(m[0], 0); M[1] = cplx(0, m[1]); M[2] = cplx(0, m[2]); M[3] = cplx(m[3], 0); break;
            case OP_SWAP: M[0] = M[3] = 0.0; M[1] = M[2] = 1.0; break;
            case OP_YL: M[0] = M[3] = 0.0; M[1] = cplx(0, -1); M[2] = cplx(0, 1); break;
            case OP_DIAG_R: M[0] = cplx(m[0], m[1]); M[1] = M[2] = 0.0; M[3] = cplx(m[6], m[7]); break;
            case OP_DIAG_R1: M[0] = 1.0; M[1] = M[2] = 0.0; M[3] = cplx(m[6], m[7]); break;
            case OP_ZFLIP: M[0] = 1.0; M[1] = M[2] = 0.0; M[3] = -1.0; break;
            default: {   // butterflies [[1, p], [q, -p q]]: the scalar alpha is already part of the launch's deferred factor
                const double* pq = HQ_BF_PQ[kind - OP_BF0];
                const cplx p(pq[0], pq[1]), q(pq[2], pq[3]);
                M[0] = 1.0; M[1] = p; M[2] = q; M[3] = -p * q;
            }
        }
    }

    void mul_matching(uint32_t creg, cplx d) {
        for (int i = 0; i < R; ++i) if ((i & creg) == creg) mul_amp(i, d);
    }

    // "multiply the amplitudes matching creg by d1 (target bit set or no target) / d0 (target bit clear), where cphys holds"
    void emit_diag_t(const DevOp& op) {
        const cplx d0(op.m[0], op.m[1]), d1(op.m[6], op.m[7]);
        const bool d0one = op.flags & 1u;
        const std::string cc = pred_all(op.cphys);
        if (op.tphys == 0) {
            if (cc.empty()) { mul_matching(op.creg, d1); return; }
            cond_begin(cc); mul_matching(op.creg, d1); cond_end();
            return;
        }
        const std::string tb = pred_all(op.tphys);
        cond_begin(both(cc, tb)); mul_matching(op.creg, d1); cond_end();
        if (!d0one && !(d0 == cplx(1.0, 0.0))) {
            cond_begin(both(cc, "!(" + tb + ")")); mul_matching(op.creg, d0); cond_end();
        }
    }

    // header + n entries: every amplitude matching header.creg is multiplied by the product of the entries that apply
    void emit_diag_run(const DevOp* hdr) {
        const int n = (int)hdr->aux;
        const uint32_t creg = hdr->creg;
        cplx S(1.0, 0.0);
        std::vector<const DevOp*> rt_entries;
        for (int e = 1; e <= n; ++e) {
            const DevOp& d = hdr[e];
            if (d.cphys == 0 && d.tphys == 0) S *= cplx(d.m[6], d.m[7]);
            else rt_entries.push_back(&d);
        }
        if (rt_entries.empty()) { if (S != cplx(1.0, 0.0)) mul_matching(creg, S); return; }
        bool all_real = true, all_sign = true;
        for (const DevOp* d : rt_entries) {
            const bool d0one = d->flags & 1u;
            if (d->m[7] != 0.0 || (!d0one && d->m[1] != 0.0)) all_real = false;
            if (std::fabs(d->m[6]) != 1.0 || (!d0one && std::fabs(d->m[0]) != 1.0)) all_sign = false;
        }
        all_sign = all_sign && all_real;
        // a real or purely imaginary static part is free (renaming); anything else starts the run-time factor
        cplx init(1.0, 0.0);
        if (S.imag() == 0.0 || S.real() == 0.0) { if (S != cplx(1.0, 0.0)) mul_matching(creg, S); }
        else { init = S; all_real = false; all_sign = false; }
        const int id = fresh();
        const std::string fr = "fr" + std::to_string(id), fi = "fi" + std::to_string(id);
        line("double " + fr + " = " + lit(init.real()) + (all_real ? ";" : ", " + fi + " = " + lit(init.imag()) + ";"));
        auto update = [&](cplx d) {
            if (d.imag() == 0.0) {
                if (d.real() == -1.0) return "{ " + fr + " = -" + fr + ";" + (all_real ? "" : " " + fi + " = -" + fi + ";") + " }";
                return "{ " + fr + " *= " + lit(d.real()) + ";" + (all_real ? "" : " " + fi + " *= " + lit(d.real()) + ";") + " }";
            }
            return "{ const double t_ = fma(" + lit(-d.imag()) + ", " + fi + ", " + fr + " * " + lit(d.real()) + "); " + fi + " = fma(" +
                   lit(d.imag()) + ", " + fr + ", " + fi + " * " + lit(d.real()) + "); " + fr + " = t_; }";
        };
        for (const DevOp* d : rt_entries) {
            const cplx d0(d->m[0], d->m[1]), d1(d->m[6], d->m[7]);
            const bool d0one = (d->flags & 1u) || d0 == cplx(1.0, 0.0);
            const std::string cc = pred_all(d->cphys);
            if (d->tphys == 0) { line("if (" + cc + ") " + update(d1)); continue; }
            const std::string tb = pred_all(d->tphys);
            if (d0one) line("if (" + both(cc, tb) + ") " + update(d1));
            else if (cc.empty()) line("if (" + tb + ") " + update(d1) + " else " + update(d0));
            else line("if (" + cc + ") { if (" + tb + ") " + update(d1) + " else " + update(d0) + " }");
        }
        if (all_sign) {   // the factor is +-1: flip signs on the integer pipe
            const std::string sg = "sg" + std::to_string(id);
            line("const int " + sg + " = __double2hiint(" + fr + ") & 0x80000000;");
            for (int i = 0; i < R; ++i) {
                if ((i & creg) != creg) continue;
                for (Term* t : {&re[i], &im[i]}) line(V(t->var) + " = hq_xs(" + V(t->var) + ", " + sg + ");");
            }
            return;
        }
        if (all_real) {
            for (int i = 0; i < R; ++i) {
                if ((i & creg) != creg) continue;
                for (Term* t : {&re[i], &im[i]}) { line(V(t->var) + " *= " + fr + ";"); ++stat_fp; }
            }
            return;
        }
        // complex factor f on c_x X + i c_y Y:  re' = c_x (fr X - (fi c_y/c_x) Y),  im' = c_y (fr Y + (fi c_x/c_y) X)
        std::map<double, std::string> scaled;   // rho -> name of fi * rho
        auto fi_times = [&](double rho) {
            if (rho == 1.0) return fi;
            auto it = scaled.find(rho);
            if (it != scaled.end()) return it->second;
            const std::string nm = "fs" + std::to_string(fresh());
            line("const double " + nm + " = " + fi + " * " + lit(rho) + ";");
            scaled[rho] = nm;
            return nm;
        };
        for (int i = 0; i < R; ++i) {
            if ((i & creg) != creg) continue;
            const Term x = re[i], y = im[i];
            const std::string a = fi_times(y.c / x.c), b = fi_times(x.c / y.c);
            const int nx = fresh(), ny = fresh();
            line("double " + V(nx) + " = fma(-" + a + ", " + V(y.var) + ", " + fr + " * " + V(x.var) + ");");
            line("double " + V(ny) + " = fma(" + b + ", " + V(x.var) + ", " + fr + " * " + V(y.var) + ");");
            stat_fp += 4;
            re[i] = Term{nx, x.c}; im[i] = Term{ny, y.c};
        }
    }

    // ---- block fusion -------------------------------------------------------------------------------------------------
    // Consecutive static ops (no run-time predicate) of a round whose register bits -- target and register-bit controls -- stay
    // within ONE or TWO register qubits form a block.  A block can be emitted gate by gate, or as the product matrix of its
    // gates (2x2 or 4x4 complex, composed here on the host) applied once: u3 ; u3 on a qubit costs 3 instead of 6 instructions
    // per scalar, the eight u3 and three cx of a quantum-volume SU(4) block 7 instead of 24, while two butterflies on two
    // qubits stay cheaper gate by gate (2 against 3).  Both forms are emitted on a trial basis and the shorter one is kept.
    // Blocks on disjoint register bits commute, so several stay open at once; anything that does not commute with a block (a
    // predicated op or a run-time diagonal on one of its bits) closes it first.
    struct Block { uint32_t bits; std::vector<const DevOp*> ops; };

    void apply_dense(const std::vector<cplx>& M, const std::vector<int>& bits) {   // M: 2^m x 2^m row-major on register bits `bits`
        const int m = (int)bits.size(), K = 1 << m;
        uint32_t bmask = 0;
        for (int b : bits) bmask |= 1u << b;
        for (int base = 0; base < R; ++base) {
            if (base & bmask) continue;
            std::vector<int> idx(K);
            for (int j = 0; j < K; ++j) {
                int v = base;
                for (int b = 0; b < m; ++b) if (j >> b & 1) v |= 1 << bits[b];
                idx[j] = v;
            }
            std::vector<Term> nr(K), ni(K);
            for (int i = 0; i < K; ++i) {
                std::vector<Term> tr, ti;
                for (int j = 0; j < K; ++j) {
                    const cplx c = M[(size_t)i * K + j];
                    tr.push_back(sc(re[idx[j]], c.real())); tr.push_back(sc(im[idx[j]], -c.imag()));
                    ti.push_back(sc(im[idx[j]], c.real())); ti.push_back(sc(re[idx[j]], c.imag()));
                }
                nr[i] = lincomb(tr);
                ni[i] = lincomb(ti);
            }
            for (int i = 0; i < K; ++i) { re[idx[i]] = nr[i]; im[idx[i]] = ni[i]; }
        }
    }

    void emit_static_op(const DevOp& op) {
        const uint32_t kind = op.code / 24, tb = (op.code % 24) / 6;
        cplx M[4];
        decode2x2(op, kind, M);
        apply2x2(M, (int)tb, op.creg);
    }

    void emit_block(const Block& blk) {
        if (blk.ops.size() == 1 || !fuse_blocks()) { for (const DevOp* op : blk.ops) emit_static_op(*op); return; }
        std::vector<int> bits;
        for (int b = 0; b < RBITS; ++b) if (blk.bits >> b & 1) bits.push_back(b);
        const int m = (int)bits.size(), K = 1 << m;
        // product of the block's gates on its own 2^m-dimensional space: apply each gate to the columns of the identity
        std::vector<cplx> U((size_t)K * K, 0.0);
        for (int i = 0; i < K; ++i) U[(size_t)i * K + i] = 1.0;
        auto local = [&](int regbit) { for (int b = 0; b < m; ++b) if (bits[b] == regbit) return b; return -1; };
        for (const DevOp* op : blk.ops) {
            const uint32_t kind = op->code / 24, tb = (op->code % 24) / 6;
            cplx G[4];
            decode2x2(*op, kind, G);
            const int t = local((int)tb);
            int cmask = 0;
            for (int b = 0; b < RBITS; ++b) if (op->creg >> b & 1) cmask |= 1 << local(b);
            for (int col = 0; col < K; ++col)
                for (int lo = 0; lo < K; ++lo) {
                    if ((lo >> t & 1) || (lo & cmask) != cmask) continue;
                    const int hi = lo | (1 << t);
                    const cplx a = U[(size_t)lo * K + col], b = U[(size_t)hi * K + col];
                    U[(size_t)lo * K + col] = G[0] * a + G[1] * b;
                    U[(size_t)hi * K + col] = G[2] * a + G[3] * b;
                }
        }
        for (cplx& v : U) {   // entries that are zero / real / imaginary up to rounding are made exactly so (they decide what is emitted)
            if (std::abs(v) < 1e-15) v = 0.0;
            else if (std::fabs(v.imag()) < 1e-15 * std::abs(v)) v = cplx(v.real(), 0.0);
            else if (std::fabs(v.real()) < 1e-15 * std::abs(v)) v = cplx(0.0, v.imag());
        }
        // trial emission of both forms; keep the one with fewer FP64 instructions
        const size_t o0 = o.size();
        const int v0 = nvar, f0 = stat_fp;
        Term r0[R], i0[R];
        std::memcpy(r0, re, sizeof(re));
        std::memcpy(i0, im, sizeof(im));
        for (const DevOp* op : blk.ops) emit_static_op(*op);
        const int eager = stat_fp - f0;
        o.resize(o0); nvar = v0; stat_fp = f0;
        std::memcpy(re, r0, sizeof(re));
        std::memcpy(im, i0, sizeof(im));
        apply_dense(U, bits);
        const int fused = stat_fp - f0;
        if (fused <= eager) { ++stat_fused; return; }
        o.resize(o0); nvar = v0; stat_fp = f0;
        std::memcpy(re, r0, sizeof(re));
        std::memcpy(im, i0, sizeof(im));
        for (const DevOp* op : blk.ops) emit_static_op(*op);
    }

    static bool fuse_blocks() { static const bool on = getenv("HQ_JIT_NO_FUSE") == nullptr; return on; }

    void emit_ops(const DevRound& rd) {
        std::vector<Block> open;
        auto close_touching = [&](uint32_t bits) {
            for (size_t i = 0; i < open.size();) {
                if (open[i].bits & bits) { emit_block(open[i]); open.erase(open.begin() + i); }
                else ++i;
            }
        };
        for (int k = rd.op_begin; k < rd.op_end; ++k) {
            const DevOp& op = ops[k];
            if (op.code == CODE_DIAG_RUN) {
                // a factor on every amplitude of the thread commutes with all static ops; one restricted to some register
                // bits is diagonal there and must follow what is pending on them
                if (op.creg) close_touching(op.creg);
                emit_diag_run(&op);
                k += (int)op.aux;
                continue;
            }
            if (op.code == CODE_DIAG_T) { if (op.creg) close_touching(op.creg); emit_diag_t(op); continue; }
            const uint32_t tb = (op.code % 24) / 6;
            const uint32_t S = (1u << tb) | op.creg;
            if (op.cphys) {   // predicated: a section of its own
                close_touching(S);
                const uint32_t kind = op.code / 24;
                cplx M[4];
                decode2x2(op, kind, M);
                cond_begin(pred_all(op.cphys)); apply2x2(M, (int)tb, op.creg); cond_end();
                continue;
            }
            uint32_t uni = S;
            for (const Block& b : open) if (b.bits & S) uni |= b.bits;
            if (__builtin_popcount(uni) > 2) {   // would outgrow a 4x4 block
                close_touching(S);
                if (__builtin_popcount(S) > 2) { emit_static_op(op); continue; }
                uni = S;
            }
            // merge the op and the open blocks it touches into one block (program order inside a block is kept by construction:
            // blocks being merged act on disjoint bits, so their relative order is free)
            Block nb;
            nb.bits = uni;
            for (size_t i = 0; i < open.size();) {
                if (open[i].bits & S) { nb.ops.insert(nb.ops.end(), open[i].ops.begin(), open[i].ops.end()); open.erase(open.begin() + i); }
                else ++i;
            }
            nb.ops.push_back(&op);
            open.push_back(std::move(nb));
        }
        close_touching(~0u);
    }

    // ---- rounds --------------------------------------------------------------------------------------------------------
    std::string stored(const Term& t, double g) {   // expression of the scalar to store, less the common factor g
        const double k = t.c / g;
        if (k == 1.0) return V(t.var);
        if (k == -1.0) return "-" + V(t.var);
        ++stat_fp;
        return "(" + lit(k) + " * " + V(t.var) + ")";
    }

    void emit_round(int r) {
        const DevRound& rd = rounds[r];
        const hq_group_plan::RoundMeta& mt = plan.meta[r];
        const bool last = rd.flags & 2u, lin_in = r == 0;
        for (int i = 0; i < 16; ++i) thread_of_tile[i] = -1;
        std::vector<std::pair<int, int>> to_tile, to_phys;
        for (size_t b = 0; b < mt.tbits.size(); ++b) {
            thread_of_tile[mt.tbits[b]] = (int)b;
            to_tile.push_back({(int)b, mt.tbits[b]});
            to_phys.push_back({(int)b, tile_phys[mt.tbits[b]]});
        }
        if (host) o += "      for (u32 tid = 0; tid < NT; ++tid)\n";   // serial replay of the CTA's threads
        o += "      {   // ---- round " + std::to_string(r) + " ----\n";
        line("const u32 tj = " + deposit("tid", to_tile, false) + ";");
        line(std::string("const u32 tin = ") + (lin_in ? "tj" : "HQ_SWZ(tj)") + ";");
        for (int i = 0; i < R; ++i) {
            const int a = fresh(), b = fresh();
            if (zero_in && r == 0)   // the only non-zero input amplitude is index 0 of tile 0 (linear position 0 of the first tile)
                line("double " + V(a) + " = (z0 && (tin ^ " + hex32(rd.ro_in[i]) + ") == 0u) ? 1.0 : 0.0, " + V(b) + " = 0.0;");
            else
            line("double " + V(a) + ", " + V(b) + "; HQ_LD(tin ^ " + hex32(rd.ro_in[i]) + ", " + V(a) + ", " + V(b) + ");");
            re[i] = Term{a, last ? deferred : 1.0};
            im[i] = Term{b, last ? deferred : 1.0};
        }
        if (last) line("HQ_TILE_CONSUMED();");
        emit_ops(rd);
        if (last) {
            line("const u64 gp = tbase | " + deposit("tid", to_phys, true) + ";");
            for (int i = 0; i < R; ++i)
                line("HQ_ST_GLOBAL(gp | " + hex64(rd.go[i]) + ", " + stored(re[i], 1.0) + ", " + stored(im[i], 1.0) + ");");
        } else {
            // the most common |coefficient| is left out of the stores and joins the launch's deferred scalar
            std::map<double, int> votes;
            for (int i = 0; i < R; ++i) { ++votes[std::fabs(re[i].c)]; ++votes[std::fabs(im[i].c)]; }
            double g = 1.0; int best = -1;
            for (auto& v : votes) if (v.second > best) { best = v.second; g = v.first; }
            deferred *= g;
            line(std::string("const u32 tout = HQ_SWZ(tj);"));
            if (rd.flags & 1u) line((rd.flags & 8u) ? "HQ_SYNCWARP();" : "HQ_SYNC();");
            for (int i = 0; i < R; ++i)
                line("HQ_ST(tout ^ " + hex32(rd.ro_out[i]) + ", " + stored(re[i], g) + ", " + stored(im[i], g) + ");");
            line((rd.flags & 4u) ? "HQ_SYNCWARP();" : "HQ_SYNC();");
            line("HQ_ROUND_DONE();");
        }
        o += "      }\n";
        if (host && !last) o += "      { double2* sw_ = cur; cur = nxt; nxt = sw_; }\n";
    }

    std::string tile_base_expr() const {   // tile number t -> base index (bits outside the tile and the fixed bits)
        const GroupParams& P = plan.p;
        std::string s = hex64(P.fixed_base);
        for (int i = 0; i < P.nseg; ++i)
            s += " | (((t >> " + std::to_string(P.seg_src[i]) + ") & " + hex64(P.seg_mask[i]) + ") << " + std::to_string(P.seg_shift[i]) + ")";
        return s;
    }

    std::string run() {
        const GroupParams& P = plan.p;
        int run_bits = 0;
        while (run_bits < K && (plan.tile_mask >> run_bits & 1)) ++run_bits;
        std::vector<std::pair<int, int>> run_dep;   // run number bit -> physical bit
        for (int j = run_bits; j < K; ++j) run_dep.push_back({j - run_bits, tile_phys[j]});
        char head[2304];
        snprintf(head, sizeof(head),
                 "// generated by hyquas_b200 group_jit: L=%d K=%d rounds=%d ops=%d gates=%d tile_mask=0x%llx\n"
                 "#define NT %d\n#define TILE %d\n#define NTILES %lluull\n#define NRUNS %d\n#define RUN_AMPS %u\n#define MINB %d\n#define PF_SLOTS %du\n",
                 plan.L, K, plan.nrounds, plan.nops, plan.ngates, (unsigned long long)plan.tile_mask, NT, TILE,
                 (unsigned long long)P.ntiles, P.nruns, P.run_bytes >> 4, jit_min_blocks(K), jit_l2_prefetch_slots());
        o = head;
        o += "#define HQ_TILE_BASE(t) (" + tile_base_expr() + ")\n";
        o += "#define HQ_RUN_OFF(q) (" + deposit("q", run_dep, true) + ")\n";
        if (zero_in) o += "#define HQ_ZERO_INPUT 1\n";
        o += host ? jit_host_prologue() : jit_device_prologue();
        for (int r = 0; r < plan.nrounds; ++r) emit_round(r);
        o += host ? jit_host_epilogue() : jit_device_epilogue();
        char tail[192];
        snprintf(tail, sizeof(tail), "// fp64 instructions per thread per tile: %d (%.2f per amplitude); %d fused blocks\n", stat_fp,
                 stat_fp / (double)R, stat_fused);
        o += tail;
        return o;
    }
};

}  // namespace

// how many slots ahead of the one being consumed a tile is pulled into L2 (<= 3: off; HQ_JIT_L2_PREFETCH)
int jit_l2_prefetch_slots() {
    static const int v = [] { const char* e = getenv("HQ_JIT_L2_PREFETCH"); return e ? std::max(0, std::min(atoi(e), 64)) : 0; }();
    return v;
}
int jit_min_blocks(int K) { return K == 12 ? 1 : (K == 11 ? 2 : 4); }   // CTAs (of two workers, three buffers) per SM
size_t jit_smem_bytes(int K) { return (size_t)3 * (16u << K) + 64; }

// Skeleton of the device kernel.  Same tile pipeline as group_kernel<K> (group_kernel.cu) with one change that ncu asked for
// (profiles/r02_s1: 38 % of all stall samples sat in the "tile landed" wait, DRAM at 4.4 TB/s): a CTA is TWO workers of NT
// threads that share THREE tile buffers.  Tiles ("slots") of a CTA are processed alternately by the two workers; slot s lives
// in buffer s % 3 and its TMA load is issued by the worker that empties that buffer (slot s - 3, when its last round has pulled
// the amplitudes into registers).  So while both workers compute, the third buffer is being filled: a load is in flight for a
// whole tile-time instead of one round.  Workers synchronise among themselves with named barriers (bar.sync 1 + worker).
// "Landed" mbarriers: SIX, slot s uses s % 6 (two per buffer, alternating).  With one per buffer the other worker can reach its
// wait for slot s + 3 while slot s is still in flight (tiny tiles, it is the one that issued slot s two short tile-times
// ago); a parity wait cannot tell "one phase ahead" from "one phase behind" and would fall through on stale data.  With two
// per buffer the barrier of slot s + 3 is idle until its own load is issued, and slot s + 6 belongs to the same worker as
// slot s, which has consumed it by then.
const char* jit_device_prologue() {
    return R"SRC(
typedef unsigned long long u64;
typedef unsigned int u32;
#define HQ_SWZ(j) ((j) ^ ((((j) >> 3) ^ ((j) >> 6) ^ ((j) >> 9) ^ ((j) >> 12)) & 7u))
__device__ __forceinline__ double hq_xs(double v, int s) { return __hiloint2double(__double2hiint(v) ^ s, __double2loint(v)); }
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// One tile = NRUNS bulk copies.  The whole worker issues them (thread q copies run q, q + NT, ...): when one warp issued all 128
// copies of a scattered tile it reached the next barrier a microsecond after the others (r02_s5: 42 gates, 7.75 ms on scattered
// tiles against 6.4 ms on contiguous ones).  Thread 0 posts the byte count; copies that complete before it did only drive the
// transaction count negative, the phase cannot complete before the arrive.
__device__ __forceinline__ void issue_tile_load(double2* state, u64 t, double2* tile, u64* bar, u32 tid, u32 nthreads) {
    const u64 base = HQ_TILE_BASE(t);
    if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(TILE * 16) : "memory");
    for (u32 q = tid; q < NRUNS; q += nthreads)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(smem_u32(tile + (size_t)q * RUN_AMPS)), "l"(state + base + HQ_RUN_OFF(q)), "r"(RUN_AMPS * 16), "r"(smem_u32(bar)) : "memory");
}
// Optional (HQ_JIT_L2_PREFETCH=n > 3, default off): pull the tile n slots ahead into L2.  Shared memory holds three tiles per SM,
// about what HBM needs in flight (6.5 TB/s x 1.5 us / 148 SMs = 66 KB); measured on supremacy_30 (r02_s4) the prefetch changes
// nothing (78.7 vs 78.8 ms): the three-buffer rotation already keeps HBM busy.
__device__ __forceinline__ void prefetch_tile_l2(const double2* state, u64 t, u32 lane) {
    const u64 base = HQ_TILE_BASE(t);
    for (u32 q = lane; q < NRUNS; q += 32)
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(state + base + HQ_RUN_OFF(q)), "r"(RUN_AMPS * 16) : "memory");
}
#define HQ_LD(idx, a, b) { const double2 q_ = tile[idx]; a = q_.x; b = q_.y; }
#define HQ_ST(idx, a, b) tile[idx] = make_double2(a, b)
#define HQ_ST_GLOBAL(idx, a, b) state[idx] = make_double2(a, b)
#define HQ_SYNC() asm volatile("bar.sync %0, %1;" :: "r"(wk + 1), "n"(NT) : "memory")
#define HQ_SYNCWARP() __syncwarp()
#define HQ_ROUND_DONE()
// every thread of the worker holds its amplitudes in registers: the buffer can take the tile three slots ahead
#ifdef HQ_ZERO_INPUT
// Input = |0...0>: there is nothing to load (and nobody zero-filled the state): each worker keeps one private tile buffer for the
// exchanges between rounds, the first round starts from constants, every tile is written as usual.
#define HQ_TILE_CONSUMED()
extern "C" __global__ void __launch_bounds__(2 * NT, MINB) hq_group_jit(double2* __restrict__ state, int amp0) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const u32 tid = threadIdx.x & (NT - 1);
    const u32 wk = threadIdx.x / NT;
    const u32 nslots = blockIdx.x < NTILES ? (u32)((NTILES - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
    double2* tile = reinterpret_cast<double2*>(smem_raw + (size_t)wk * TILE * 16);
    for (u32 s = wk; s < nslots; s += 2) {
      const u64 tbase = HQ_TILE_BASE((u64)blockIdx.x + (u64)s * gridDim.x);
      const bool z0 = amp0 != 0 && tbase == 0;
      if (s >= 2) HQ_SYNC();   // the previous tile's last round has read this buffer
#else
#define HQ_TILE_CONSUMED() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); HQ_SYNC(); \
        if (s + 3 < nslots) issue_tile_load(state, (u64)blockIdx.x + (u64)(s + 3) * gridDim.x, tile, bar + (s + 3) % 6u, tid, NT); \
        if (PF_SLOTS > 3 && s + PF_SLOTS < nslots && tid < 32) prefetch_tile_l2(state, (u64)blockIdx.x + (u64)(s + PF_SLOTS) * gridDim.x, tid); }
extern "C" __global__ void __launch_bounds__(2 * NT, MINB) hq_group_jit(double2* __restrict__ state, int amp0) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    u64* bar = reinterpret_cast<u64*>(smem_raw + (size_t)3 * TILE * 16);
    const u32 tid = threadIdx.x & (NT - 1);
    const u32 wk = threadIdx.x / NT;
    const u32 nslots = blockIdx.x < NTILES ? (u32)((NTILES - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;   // tiles of this CTA
    if (threadIdx.x == 0) {
        for (int i = 0; i < 6; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar + i)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    for (u32 i = 0; i < 3 && i < nslots; ++i)
        issue_tile_load(state, (u64)blockIdx.x + (u64)i * gridDim.x, reinterpret_cast<double2*>(smem_raw + (size_t)i * TILE * 16), bar + i, threadIdx.x, 2 * NT);
    for (u32 s = wk; s < nslots; s += 2) {
      const u32 b = s % 3u;
      double2* tile = reinterpret_cast<double2*>(smem_raw + (size_t)b * TILE * 16);
      mbar_wait(bar + s % 6u, (s / 6u) & 1u);
      const u64 tbase = HQ_TILE_BASE((u64)blockIdx.x + (u64)s * gridDim.x);
#endif
)SRC";
}
const char* jit_device_epilogue() { return "    }\n}\n"; }

// Host flavour (test infrastructure): tiles and threads replayed serially, reads of a round before its writes.
const char* jit_host_prologue() {
    return R"SRC(
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
typedef unsigned long long u64;
typedef unsigned int u32;
struct double2 { double x, y; };
#define HQ_SWZ(j) ((j) ^ ((((j) >> 3) ^ ((j) >> 6) ^ ((j) >> 9) ^ ((j) >> 12)) & 7u))
static inline int __double2hiint(double v) { u64 b; std::memcpy(&b, &v, 8); return (int)(b >> 32); }
static inline double hq_xs(double v, int s) { u64 b; std::memcpy(&b, &v, 8); b ^= (u64)(u32)s << 32; std::memcpy(&v, &b, 8); return v; }
#define HQ_LD(idx, a, b) { const double2 q_ = cur[idx]; a = q_.x; b = q_.y; }
#define HQ_ST(idx, a, b) nxt[idx] = double2{a, b}
#define HQ_ST_GLOBAL(idx, a, b) state[idx] = double2{a, b}
#define HQ_SYNC()
#define HQ_SYNCWARP()
#define HQ_TILE_CONSUMED()
#define HQ_ROUND_DONE()
// one function per round would be tidier, but the arithmetic text must be the device text: a round is a block scope that the
// host skeleton runs once per thread id through this macro pair
extern "C" void hq_group_jit_host(double* state_re_im, int amp0) {
    double2* state = reinterpret_cast<double2*>(state_re_im);
    std::vector<double2> bufA(TILE), bufB(TILE);
    for (u64 t = 0; t < NTILES; ++t) {
      const u64 tbase = HQ_TILE_BASE(t);
      const bool z0 = amp0 != 0 && tbase == 0; (void)z0;
      double2* cur = bufA.data(); double2* nxt = bufB.data();
#ifndef HQ_ZERO_INPUT
      for (u32 q = 0; q < NRUNS; ++q) std::memcpy(cur + (size_t)q * RUN_AMPS, state + tbase + HQ_RUN_OFF(q), (size_t)RUN_AMPS * 16);
#endif
)SRC";
}
const char* jit_host_epilogue() { return "    }\n}\n"; }

std::string jit_emit_source(const hq_group_plan& plan, bool host, bool zero_input) {
    Emitter e(plan, host);
    e.zero_in = zero_input;
    std::string src = e.run();
    if (!e.ok) return std::string();
    return src;
}

// FP64 instructions per amplitude of the specialised kernel (what the emitter would write), without keeping the source
double jit_fp64_per_amp(const hq_group_plan& plan) {
    Emitter e(plan, false);
    e.run();
    return e.ok ? e.stat_fp / (double)R : -1.0;
}

}  // namespace hq

extern "C" int hq_group_plan_cost(const hq_group_plan* plan, int* rounds, double* fp64_per_amp) {
    HQ_REQUIRE(plan != nullptr, "null plan");
    if (rounds) *rounds = plan->nrounds;
    if (fp64_per_amp) *fp64_per_amp = hq::jit_fp64_per_amp(*plan);
    return HQ_OK;
}

// host_flavour: 0 = CUDA source, 1 = host source; +2 = the zero-input variant (input |0...0>, nothing loaded)
extern "C" int hq_debug_group_plan_jit_source(const hq_group_plan* plan, int host_flavour, char* out, size_t cap, size_t* needed) {
    HQ_REQUIRE(plan != nullptr && needed != nullptr, "null argument");
    const std::string src = hq::jit_emit_source(*plan, (host_flavour & 1) != 0, (host_flavour & 2) != 0);
    HQ_REQUIRE(!src.empty(), "the JIT emitter rejected this plan");
    *needed = src.size() + 1;
    if (out && cap >= src.size() + 1) std::memcpy(out, src.c_str(), src.size() + 1);
    return HQ_OK;
}
