#!/usr/bin/env python
"""Generates hyquas_b200/csrc/device/group_ops_gen.inc: the in-register gate arithmetic of the gate-group kernel.

Why generated inline PTX: the 16 amplitudes a thread holds (32 FP64 values) must stay in the SAME registers across the
op loop and its 100+ way switch.  Written as a C++ array, LLVM turns them into 64 SSA webs with phi nodes at every
join and ptxas ends up copying the whole set once per gate (measured: 65 % of all executed instructions were MOVs).
Here the amplitudes live in named PTX registers (hqa0..hqa31) that the C++ compiler never sees; every op body
updates them in place, so a gate costs its FP64 instructions and nothing else.

    python tools/gen_group_ops.py 4 > hyquas_b200/csrc/device/group_ops_gen_r4.inc     (argument = register qubits per thread)

Register naming: amplitude i (0..15) = (hqa{2i}, hqa{2i+1}) = (re, im).
Op numbering must match group_plan.h: code = kind*24 + tb*6 + cbc, cbc: 0 none, 1..4 register bit cbc-1, 5 generic.
"""
import sys

RBITS = int(sys.argv[1]) if len(sys.argv) > 1 else 4
R = 1 << RBITS
KINDS = ["GEN", "REAL", "RXL", "SWAP", "YL", "DIAG_R", "ZFLIP"]
GENERIC_ONLY = {"GEN", "REAL", "RXL", "YL"}   # single-control cases are routed to the generic mask on the host


def re(i):
    return f"hqa{2 * i}"


def im(i):
    return f"hqa{2 * i + 1}"


def pairs(tb):
    out = []
    for p in range(R // 2):
        lo = ((p >> tb) << (tb + 1)) | (p & ((1 << tb) - 1))
        out.append((lo, lo | (1 << tb)))
    return out


def lu2(u, v, c, d, e, f):
    """v <- c*u + d*v ; u <- e*u + f*v   (in place)"""
    return [f"mul.f64 {v}, {d}, {v};", f"fma.rn.f64 {v}, {c}, {u}, {v};", f"mul.f64 {u}, {e}, {u};",
            f"fma.rn.f64 {u}, {f}, {v}, {u};"]


def caxpby(zx, zy, wx, wy, pr, pi, npi, qr, qi, nqi, t="hqt"):
    """z <- p*z + q*w, one temporary (npi / nqi are the negated imaginary parts)"""
    return [f"mul.f64 {t}, {pr}, {zx};", f"fma.rn.f64 {t}, {npi}, {zy}, {t};", f"fma.rn.f64 {t}, {qr}, {wx}, {t};",
            f"fma.rn.f64 {t}, {nqi}, {wy}, {t};",
            f"mul.f64 {zy}, {pr}, {zy};", f"fma.rn.f64 {zy}, {pi}, {zx}, {zy};", f"fma.rn.f64 {zy}, {qr}, {wy}, {zy};",
            f"fma.rn.f64 {zy}, {qi}, {wx}, {zy};", f"mov.f64 {zx}, {t};"]


def cmul(zx, zy, pr, pi, npi, t="hqt"):
    return [f"mul.f64 {t}, {pr}, {zx};", f"fma.rn.f64 {t}, {npi}, {zy}, {t};", f"mul.f64 {zy}, {pr}, {zy};",
            f"fma.rn.f64 {zy}, {pi}, {zx}, {zy};", f"mov.f64 {zx}, {t};"]


NEG = "0x8000000000000000"


def pair_body(kind, lo, hi):
    """PTX for one (lo, hi) pair.  Returns (lines, operand names in order, needs_temp)."""
    if kind == "REAL":      # m = {c, d, e, f}
        ops = ["c", "d", "e", "f"]
        ls = lu2(re(lo), re(hi), "%0", "%1", "%2", "%3") + lu2(im(lo), im(hi), "%0", "%1", "%2", "%3")
        return ls, ops
    if kind == "RXL":       # (lo.x, hi.y) with {c,d,e,f}; (lo.y, hi.x) with {-c,d,e,-f}  -> operands c,d,e,f,nc,nf
        ops = ["c", "d", "e", "f", "nc", "nf"]
        ls = lu2(re(lo), im(hi), "%0", "%1", "%2", "%3") + lu2(im(lo), re(hi), "%4", "%1", "%2", "%5")
        return ls, ops
    if kind == "GEN":       # hi <- c*lo + d*hi ; lo <- e*lo + f*hi
        ops = ["cr", "ci", "dr", "di", "er", "ei", "fr", "fi", "nci", "ndi", "nei", "nfi"]
        ls = caxpby(re(hi), im(hi), re(lo), im(lo), "%2", "%3", "%9", "%0", "%1", "%8")
        ls += caxpby(re(lo), im(lo), re(hi), im(hi), "%4", "%5", "%10", "%6", "%7", "%11")
        return ls, ops
    if kind == "SWAP":
        ls = []
        for x, y in ((re(lo), re(hi)), (im(lo), im(hi))):
            ls += [f"mov.f64 hqt, {x};", f"mov.f64 {x}, {y};", f"mov.f64 {y}, hqt;"]
        return ls, []
    if kind == "YL":        # lo' = -i*hi = (hi.y, -hi.x) ; hi' = i*lo = (-lo.y, lo.x)
        ls = [f"mov.f64 hqt, {re(lo)};", f"mov.f64 hqu, {im(lo)};",
              f"mov.f64 {re(lo)}, {im(hi)};", f"xor.b64 {im(lo)}, {re(hi)}, {NEG};",
              f"xor.b64 {re(hi)}, hqu, {NEG};", f"mov.f64 {im(hi)}, hqt;"]
        return ls, []
    if kind == "DIAG_R1":   # hi *= d1
        return cmul(re(hi), im(hi), "%0", "%1", "%2"), ["r1", "i1", "ni1"]
    if kind == "DIAG_R01":  # lo *= d0, hi *= d1
        return (cmul(re(lo), im(lo), "%3", "%4", "%5") + cmul(re(hi), im(hi), "%0", "%1", "%2"),
                ["r1", "i1", "ni1", "r0", "i0", "ni0"])
    if kind == "ZFLIP":
        return [f"xor.b64 {re(hi)}, {re(hi)}, {NEG};", f"xor.b64 {im(hi)}, {im(hi)}, {NEG};"], []
    raise ValueError(kind)


OPERAND_EXPR = {
    "REAL": {"c": "o.m[0]", "d": "o.m[1]", "e": "o.m[2]", "f": "o.m[3]"},
    "RXL": {"c": "o.m[0]", "d": "o.m[1]", "e": "o.m[2]", "f": "o.m[3]", "nc": "-o.m[0]", "nf": "-o.m[3]"},
    "GEN": {"cr": "o.m[0]", "ci": "o.m[1]", "dr": "o.m[2]", "di": "o.m[3]", "er": "o.m[4]", "ei": "o.m[5]",
            "fr": "o.m[6]", "fi": "o.m[7]", "nci": "-o.m[1]", "ndi": "-o.m[3]", "nei": "-o.m[5]", "nfi": "-o.m[7]"},
    "DIAG_R1": {"r1": "o.m[6]", "i1": "o.m[7]", "ni1": "-o.m[7]"},
    "DIAG_R01": {"r1": "o.m[6]", "i1": "o.m[7]", "ni1": "-o.m[7]", "r0": "o.m[0]", "i0": "o.m[1]", "ni0": "-o.m[1]"},
}


def asm_stmt(lines, operands, kind, indent="    "):
    body = " ".join(lines)
    text = '"{ .reg .f64 hqt, hqu; ' + body + ' }"'
    if operands:
        ins = ", ".join(f'"d"({name})' for name in operands)
        return f"{indent}asm volatile({text} :: {ins});"
    return f"{indent}asm volatile({text});"


def emit_body(kind, tb, cbc):
    """C++ statements applying `kind` on target register bit tb with control case cbc."""
    out = []
    variants = [kind]
    if kind == "DIAG_R":
        variants = ["DIAG_R1", "DIAG_R01"]
    for vi, var in enumerate(variants):
        decl = OPERAND_EXPR.get(var, {})
        names = None
        stmts = []
        sel = []
        for lo, hi in pairs(tb):
            if 1 <= cbc <= RBITS and not (lo >> (cbc - 1)) & 1:
                continue
            ls, names = pair_body(var, lo, hi)
            sel.append((lo, ls))
        pre = [f"    const double {n} = {decl[n]};" for n in (names or [])]
        if cbc == 5:
            for lo, ls in sel:
                stmts.append(f"    if (({lo}u & creg) == creg) {{")
                stmts.append(asm_stmt(ls, names, var, indent="        "))
                stmts.append("    }")
        else:
            # a few pairs per asm statement keeps the strings readable and lets ptxas interleave freely
            flat = []
            for _, ls in sel:
                flat += ls
            stmts.append(asm_stmt(flat, names, var))
        block = pre + stmts
        if kind == "DIAG_R":
            cond = "if (o.flags & 1u) {" if vi == 0 else "} else {"
            out.append("    " + cond)
            out += ["    " + b for b in block]
            if vi == 1:
                out.append("    }")
        else:
            out += block
    return out


def main():
    print("// GENERATED by tools/gen_group_ops.py -- do not edit.  See that script for the why and the register naming.")
    print("// clang-format off")
    print(f'#define HQ_DECLARE_AMP_REGS() asm volatile(".reg .f64 hqa<{2 * R}>;")')
    print()
    # loads / stores
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
    # complex multiply of all / masked amplitudes by a runtime factor
    print("__device__ __forceinline__ void hq_cmul_all(double fr, double fi) {")
    print("    const double nfi = -fi;")
    ls = []
    for i in range(R):
        ls += cmul(re(i), im(i), "%0", "%1", "%2")
    print(asm_stmt(ls, ["fr", "fi", "nfi"], "cmul"))
    print("}")
    print("__device__ __forceinline__ void hq_cmul_bit(double fr, double fi, int bit) {   // amplitudes whose register-index bit `bit` is 1")
    print("    const double nfi = -fi;")
    print("    switch (bit) {")
    for b in range(RBITS):
        ls = []
        for i in range(R):
            if i >> b & 1:
                ls += cmul(re(i), im(i), "%0", "%1", "%2")
        print(f"        case {b}:")
        print(asm_stmt(ls, ["fr", "fi", "nfi"], "cmul", indent="            "))
        print("            break;")
    print("        default: break;")
    print("    }")
    print("}")
    print("__device__ __forceinline__ void hq_cmul_masked(double fr, double fi, uint32_t creg) {")
    print("    const double nfi = -fi;")
    for i in range(R):
        print(f"    if (({i}u & creg) == creg) {{")
        print(asm_stmt(cmul(re(i), im(i), "%0", "%1", "%2"), ["fr", "fi", "nfi"], "cmul", indent="        "))
        print("    }")
    print("}")
    print()
    cases = []
    for k, kind in enumerate(KINDS):
        for tb in range(RBITS):
            for cbc in range(6):
                if 1 <= cbc <= 4 and (cbc - 1 == tb or cbc > RBITS or kind in GENERIC_ONLY):
                    continue
                code = k * 24 + tb * 6 + cbc
                print(f"__device__ __forceinline__ void hq_op_{code}(const hq::DevOp& o) {{   // {kind} tb={tb} cbc={cbc}")
                if cbc == 5:
                    print("    const uint32_t creg = o.creg;")
                for l in emit_body(kind, tb, cbc):
                    print(l)
                print("}")
                cases.append(code)
    print()
    # dense body index: a contiguous 0..N-1 switch compiles to ONE indexed branch (the sparse switch on `code` became a
    # tree of compares + small BRX tables, ~8 dependent branches per gate).  The planner stores the index in flags[15:8].
    print(f"#define HQ_OP_BODIES {len(cases)}")
    table = [255] * (max(cases) + 1)
    for i, c in enumerate(cases):
        table[c] = i
    print("static const unsigned char HQ_OP_BODY_INDEX[] = {" + ", ".join(str(v) for v in table) + "};")
    print("__device__ __forceinline__ void hq_apply_op(const hq::DevOp& o) {")
    print("    switch ((o.flags >> 8) & 0xffu) {")
    for i, c in enumerate(cases):
        print(f"        case {i}: hq_op_{c}(o); break;")
    print("        default: break;")
    print("    }")
    print("}")
    print("// clang-format on")


if __name__ == "__main__":
    main()
