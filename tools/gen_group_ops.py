#!/usr/bin/env python
"""Generates hyquas_b200/csrc/device/group_ops_gen_r{3,4}.inc: the in-register gate arithmetic of the gate-group kernel.

Why generated inline PTX: the 16 amplitudes a thread holds (32 FP64 values) must stay in the SAME registers across the
op loop and its 100+ way dispatch.  Written as a C++ array, LLVM turns them into 64 SSA webs with phi nodes at every
join and ptxas ends up copying the whole set once per gate (measured: 65 % of all executed instructions were MOVs).
Here the amplitudes live in named PTX registers (hqa0..hqa31) that the C++ compiler never sees; every op body
updates them in place, so a gate costs its FP64 instructions and nothing else.

All op bodies live in ONE asm statement behind a single `brx.idx` jump table (a C++ switch over the same bodies compiled
to a 7-deep tree of compare+branch per gate).  Operands of that statement are fixed:
    %0 = body index, %1 = register-index control mask (creg), %2 = shared-memory address of the op's coefficients m[0..7].
Each body loads the coefficients it reads (none for butterflies, sign flips, swaps, Y) into registers that live only inside
the body: the op list is the kernel's largest shared-memory consumer (ncu), and coefficient registers that are live around
the op loop cost spills at the 128-register cap.

    python tools/gen_group_ops.py 4 > hyquas_b200/csrc/device/group_ops_gen_r4.inc     (argument = register qubits per thread)

Register naming: amplitude i (0..15) = (hqa{2i}, hqa{2i+1}) = (re, im).
Op numbering must match group_plan.h: code = kind*24 + tb*6 + cbc, cbc: 0 none, 1..4 register bit cbc-1, 5 generic mask.

tests/test_group_ops_ptx.py interprets the generated PTX of every body on random data and checks it against the 2x2
matrix the planner means by it, so the arithmetic is proven on a machine without a GPU.
"""
import sys

# must match enum OpKind in group_plan.h
KINDS = ["GEN", "REAL", "RXL", "SWAP", "YL", "DIAG_R", "ZFLIP", "DIAG_R1",
         "BF0", "BF1", "BF2", "BF3", "BF4", "BF5", "BF6", "BF7"]
GENERIC_ONLY = {"GEN", "REAL", "RXL", "YL"}   # single-control cases are routed to the generic mask on the host
NO_CONTROL = {k for k in KINDS if k.startswith("BF")}   # butterflies defer their scale: uncontrolled only
# butterfly variants: M = alpha * [[1, p], [q, -p*q]], (p, q) both real or both imaginary units
BF_PQ = [(1, 1), (1, -1), (-1, 1), (-1, -1), (1j, 1j), (1j, -1j), (-1j, 1j), (-1j, -1j)]
M = {f"m{i}": f"hqm{i}" for i in range(8)}
COEFF_ADDR = "%2"
CREG = "%1"
NEG = "0x8000000000000000"
TWO = "0d4000000000000000"


def re(i):
    return f"hqa{2 * i}"


def im(i):
    return f"hqa{2 * i + 1}"


def pairs(tb, R):
    out = []
    for p in range(R // 2):
        lo = ((p >> tb) << (tb + 1)) | (p & ((1 << tb) - 1))
        out.append((lo, lo | (1 << tb)))
    return out


def lin2(u, v, a, b, c, d, t="hqt"):
    """(u, v) <- (a*u + b*v, c*u + d*v) on scalars: four FP64 instructions and one register copy.  (An in-place LU form
    needs no copy but divides by d: gates with a small diagonal -- RX(theta) near pi, a third of random U3 -- then had to run
    as X * (X M), a complex general op plus a register swap, 2.5x the cost.)"""
    return [f"mov.f64 {t}, {u};", f"mul.f64 {u}, {a}, {u};", f"fma.rn.f64 {u}, {b}, {v}, {u};", f"mul.f64 {v}, {d}, {v};",
            f"fma.rn.f64 {v}, {c}, {t}, {v};"]


def caxpby(zx, zy, wx, wy, pr, pi, npi, qr, qi, nqi, t="hqt"):
    """z <- p*z + q*w, one temporary (npi / nqi are the negated imaginary parts)"""
    return [f"mul.f64 {t}, {pr}, {zx};", f"fma.rn.f64 {t}, {npi}, {zy}, {t};", f"fma.rn.f64 {t}, {qr}, {wx}, {t};",
            f"fma.rn.f64 {t}, {nqi}, {wy}, {t};",
            f"mul.f64 {zy}, {pr}, {zy};", f"fma.rn.f64 {zy}, {pi}, {zx}, {zy};", f"fma.rn.f64 {zy}, {qr}, {wy}, {zy};",
            f"fma.rn.f64 {zy}, {qi}, {wx}, {zy};", f"mov.f64 {zx}, {t};"]


def cmul(zx, zy, pr, pi, npi, t="hqt"):
    return [f"mul.f64 {t}, {pr}, {zx};", f"fma.rn.f64 {t}, {npi}, {zy}, {t};", f"mul.f64 {zy}, {pr}, {zy};",
            f"fma.rn.f64 {zy}, {pi}, {zx}, {zy};", f"mov.f64 {zx}, {t};"]


def bfly(u, v, sigma, tau):
    """u' = u + sigma*v ; v' = tau*(u - sigma*v), two FP64 instructions, in place (u' = 2u - tau*v')."""
    if tau > 0:
        first = [f"sub.f64 {v}, {u}, {v};"] if sigma > 0 else [f"add.f64 {v}, {u}, {v};"]
        return first + [f"neg.f64 hqt, {v};", f"fma.rn.f64 {u}, {u}, {TWO}, hqt;"]
    first = [f"sub.f64 {v}, {v}, {u};"] if sigma > 0 else [f"neg.f64 hqt, {u};", f"sub.f64 {v}, hqt, {v};"]
    return first + [f"fma.rn.f64 {u}, {u}, {TWO}, {v};"]


def bfly_pairs(p, q, lo, hi):
    """Scalar butterflies of M = [[1,p],[q,-pq]] on amplitudes lo=(a,b), hi=(c,d): list of (u, v, sigma, tau).

    lo' = lo + w, hi' = q*(lo - w) with w = p*hi.  For real p the scalar pairs are (a,c),(b,d); for imaginary p they are
    (a,d),(b,c), and an imaginary q puts q*(lo - w) back into the natural registers with a sign."""
    a, b, c, d = re(lo), im(lo), re(hi), im(hi)
    if p.imag == 0:
        s = 1 if p.real > 0 else -1            # w = (s c, s d)
        t = 1 if q.real > 0 else -1            # hi' = t * (a - s c, b - s d)
        assert q.imag == 0
        return [(a, c, s, t), (b, d, s, t)]
    s = 1 if p.imag > 0 else -1                # w = s*i*(c + i d) = (-s d, s c)
    t = 1 if q.imag > 0 else -1                # hi' = t*i*(x + i y) = (-t y, t x) with x = a + s d (lands in d), y = b - s c (in c)
    assert q.real == 0
    # register d: u = a, v = d: u' = a - s d ... careful: lo'.re = a + w.re = a - s d ; x = a - w.re = a + s d
    # so for (a, d): sigma = -s, and d' = hi'.im = t * x = t * (a + s d) = tau*(u - sigma v) with tau = t
    # for (b, c): lo'.im = b + w.im = b + s c -> sigma = s ; y = b - s c ; c' = hi'.re = -t * y -> tau = -t
    return [(a, d, -s, t), (b, c, s, -t)]


def pair_body(kind, lo, hi):
    """PTX lines for one (lo, hi) pair; coefficients are the statement operands m0..m7 (and negations hqn*)."""
    if kind == "REAL":      # m = {a, b, c, d}: the real matrix [[a, b], [c, d]] on (lo.x, hi.x) and on (lo.y, hi.y)
        return lin2(re(lo), re(hi), M["m0"], M["m1"], M["m2"], M["m3"]) + lin2(im(lo), im(hi), M["m0"], M["m1"], M["m2"], M["m3"])
    if kind == "RXL":       # m = {a, b, c, d} of [[a, i b], [i c, d]]: (lo.x, hi.y) sees [[a, -b], [c, d]], (lo.y, hi.x) sees [[a, b], [-c, d]]
        return lin2(re(lo), im(hi), M["m0"], "hqn1", M["m2"], M["m3"]) + lin2(im(lo), re(hi), M["m0"], M["m1"], "hqn2", M["m3"])
    if kind == "GEN":       # m = the complex matrix [[a, b], [c, d]] itself: lo' = a lo + b hi ; hi' = c lo(old) + d hi
        ls = [f"mov.f64 hqs0, {re(lo)};", f"mov.f64 hqs1, {im(lo)};"]
        ls += caxpby(re(lo), im(lo), re(hi), im(hi), M["m0"], M["m1"], "hqn1", M["m2"], M["m3"], "hqn3")
        ls += caxpby(re(hi), im(hi), "hqs0", "hqs1", M["m6"], M["m7"], "hqn7", M["m4"], M["m5"], "hqn5")
        return ls
    if kind == "SWAP":
        ls = []
        for x, y in ((re(lo), re(hi)), (im(lo), im(hi))):
            ls += [f"mov.f64 hqt, {x};", f"mov.f64 {x}, {y};", f"mov.f64 {y}, hqt;"]
        return ls
    if kind == "YL":        # lo' = -i*hi = (hi.y, -hi.x) ; hi' = i*lo = (-lo.y, lo.x)
        return [f"mov.f64 hqt, {re(lo)};", f"mov.f64 hqu, {im(lo)};",
                f"mov.f64 {re(lo)}, {im(hi)};", f"xor.b64 {im(lo)}, {re(hi)}, {NEG};",
                f"xor.b64 {re(hi)}, hqu, {NEG};", f"mov.f64 {im(hi)}, hqt;"]
    if kind == "DIAG_R1":   # hi *= d1 = (m6, m7)
        return cmul(re(hi), im(hi), M["m6"], M["m7"], "hqn7")
    if kind == "DIAG_R":    # lo *= d0 = (m0, m1), hi *= d1
        return cmul(re(lo), im(lo), M["m0"], M["m1"], "hqn1") + cmul(re(hi), im(hi), M["m6"], M["m7"], "hqn7")
    if kind == "ZFLIP":
        return [f"xor.b64 {re(hi)}, {re(hi)}, {NEG};", f"xor.b64 {im(hi)}, {im(hi)}, {NEG};"]
    if kind.startswith("BF"):
        p, q = BF_PQ[int(kind[2:])]
        ls = []
        for u, v, s, t in bfly_pairs(complex(p), complex(q), lo, hi):
            ls += bfly(u, v, s, t)
        return ls
    raise ValueError(kind)


NEGS_NEEDED = {"RXL": [1, 2], "GEN": [1, 3, 5, 7], "DIAG_R1": [7], "DIAG_R": [1, 7]}
# coefficient pairs (m[2j], m[2j+1]) a body reads: one 16-byte shared load each
COEFF_PAIRS = {"GEN": [0, 1, 2, 3], "REAL": [0, 1], "RXL": [0, 1], "DIAG_R": [0, 3], "DIAG_R1": [3]}


def body_lines(kind, tb, cbc, rbits):
    """Straight-line PTX of one dispatch target: `kind` on target register bit tb with control case cbc."""
    R = 1 << rbits
    out = [f"ld.shared.v2.f64 {{hqm{2 * j}, hqm{2 * j + 1}}}, [{COEFF_ADDR}+{16 * j}];" for j in COEFF_PAIRS.get(kind, [])]
    out += [f"neg.f64 hqn{i}, {M[f'm{i}']};" for i in NEGS_NEEDED.get(kind, [])]
    for lo, hi in pairs(tb, R):
        if 1 <= cbc <= 4 and not (lo >> (cbc - 1)) & 1:
            continue
        ls = pair_body(kind, lo, hi)
        if cbc == 5:   # generic register-control mask: pair participates iff (lo & creg) == creg
            out += [f"and.b32 hqx, {CREG}, {(~lo) & (R - 1)};", "setp.eq.u32 hqp, hqx, 0;"]
            out += ["@hqp " + l for l in ls]
        else:
            out += ls
    return out


def catalog(rbits):
    """[(code, kind, tb, cbc, lines)] in body-index order."""
    out = []
    for k, kind in enumerate(KINDS):
        for tb in range(rbits):
            for cbc in range(6):
                if 1 <= cbc <= 4 and (cbc - 1 == tb or cbc > rbits or kind in GENERIC_ONLY):
                    continue
                if cbc != 0 and kind in NO_CONTROL:
                    continue
                out.append((k * 24 + tb * 6 + cbc, kind, tb, cbc, body_lines(kind, tb, cbc, rbits)))
    return out


def asm_stmt(lines, operands, indent="    "):
    body = " ".join(lines)
    text = '"{ .reg .f64 hqt; ' + body + ' }"'
    ins = ", ".join(f'"d"({name})' for name in operands)
    return f"{indent}asm volatile({text} :: {ins});"


def main():
    rbits = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    R = 1 << rbits
    print("// GENERATED by tools/gen_group_ops.py -- do not edit.  See that script for the why and the register naming.")
    print("// clang-format off")
    print(f'#define HQ_DECLARE_AMP_REGS() asm volatile(".reg .f64 hqa<{2 * R}>;")')
    print()
    print("// tile (shared memory, 32-bit shared address of amplitude index 0) <-> amplitude registers")
    print("__device__ __forceinline__ void hq_load_amps(uint32_t tile_s, uint32_t tin, const uint16_t* ro) {")
    for i in range(R):
        print(f'    asm volatile("ld.shared.v2.f64 {{{re(i)}, {im(i)}}}, [%0];" :: "r"(tile_s + ((tin ^ ro[{i}]) << 4)) : "memory");')
    print("}")
    print("__device__ __forceinline__ void hq_store_amps(uint32_t tile_s, uint32_t tout, const uint16_t* ro) {")
    for i in range(R):
        print(f'    asm volatile("st.shared.v2.f64 [%0], {{{re(i)}, {im(i)}}};" :: "r"(tile_s + ((tout ^ ro[{i}]) << 4)) : "memory");')
    print("}")
    print("__device__ __forceinline__ void hq_store_amps_global(double2* base, const uint64_t* go) {")
    for i in range(R):
        print(f'    asm volatile("st.global.v2.f64 [%0], {{{re(i)}, {im(i)}}};" :: "l"(base + go[{i}]) : "memory");')
    print("}")
    print()

    def cm(z, fr="%0", fi="%1", nfi="%2"):
        return cmul(re(z), im(z), fr, fi, nfi)

    print("__device__ __forceinline__ void hq_cmul_all(double fr, double fi) {")
    print("    const double nfi = -fi;")
    ls = []
    for i in range(R):
        ls += cm(i)
    print(asm_stmt(ls, ["fr", "fi", "nfi"]))
    print("}")
    print("__device__ __forceinline__ void hq_cmul_bit(double fr, double fi, int bit) {   // amplitudes whose register-index bit `bit` is 1")
    print("    const double nfi = -fi;")
    print("    switch (bit) {")
    for b in range(rbits):
        ls = []
        for i in range(R):
            if i >> b & 1:
                ls += cm(i)
        print(f"        case {b}:")
        print(asm_stmt(ls, ["fr", "fi", "nfi"], indent="            "))
        print("            break;")
    print("        default: break;")
    print("    }")
    print("}")
    print("__device__ __forceinline__ void hq_cmul_masked(double fr, double fi, uint32_t creg) {")
    print("    const double nfi = -fi;")
    for i in range(R):
        print(f"    if (({i}u & creg) == creg) {{")
        print(asm_stmt(cm(i), ["fr", "fi", "nfi"], indent="        "))
        print("    }")
    print("}")
    print()

    cat = catalog(rbits)
    print(f"#define HQ_OP_BODIES {len(cat)}")
    table = [255] * (len(KINDS) * 24)
    for i, (code, *_rest) in enumerate(cat):
        table[code] = i
    assert len(cat) < 255
    print("static const unsigned char HQ_OP_BODY_INDEX[] = {" + ", ".join(str(v) for v in table) + "};")
    print("// one jump, one body: `body` is HQ_OP_BODY_INDEX[code], warp-uniform")
    print("__device__ __forceinline__ void hq_apply_op(uint32_t body, uint32_t creg, uint32_t coeff_addr) {")
    print("    asm volatile(\"{\\n\"")
    print('        ".reg .f64 hqt, hqu, hqs0, hqs1, hqn1, hqn2, hqn3, hqn5, hqn7, hqm<8>;\\n"')
    print('        ".reg .pred hqp;\\n"')
    print('        ".reg .b32 hqx;\\n"')
    labels = ", ".join(f"HQB{i}" for i in range(len(cat)))
    print(f'        "HQTS: .branchtargets {labels};\\n"')
    print('        "brx.idx.uni %0, HQTS;\\n"')
    for i, (code, kind, tb, cbc, lines) in enumerate(cat):
        print(f'        "HQB{i}:\\n"   // {kind} tb={tb} cbc={cbc} (code {code})')
        for l in lines:
            print(f'        "{l}\\n"')
        print('        "bra.uni HQEND;\\n"')
    print('        "HQEND:\\n"')
    print('        "}" :: "r"(body), "r"(creg), "r"(coeff_addr) : "memory");')
    print("}")
    print("// clang-format on")


if __name__ == "__main__":
    main()
