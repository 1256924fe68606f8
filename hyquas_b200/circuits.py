"""Circuit generators in the input dialect accepted by the QASM-subset parser
(reference grammar: main.cpp:76-223 -- one statement per line, operands without spaces,
parameters written `pi*x`, `pi/x` or a plain decimal).

The reference ships its benchmark circuits as Git-LFS payloads that are not in the checkout;
`qft`, `bv` and `hidden_shift` below reproduce the originals byte for byte (sha256 == LFS oid, see
tests/golden/lfs_oids.json), the other families are seeded synthetic instances of the same
benchmark classes (supremacy, quantum_volume, qaoa, adder, basis_change).
"""
from __future__ import annotations

import math
import random
from typing import Callable, Dict, List

HEADER = 'OPENQASM 2.0;\ninclude "qelib1.inc";\nqreg q[{n}];\n'


def _emit(n: int, lines: List[str]) -> str:
    return HEADER.format(n=n) + "".join(l + "\n" for l in lines)


def qft(n: int) -> str:
    """h q[i]; then cu1(pi/2^(j-i)) q[i],q[j] for j>i.  No final swaps (matches tests/input/qft_28.qasm)."""
    out = []
    for i in range(n):
        out.append(f"h q[{i}];")
        for j in range(i + 1, n):
            out.append(f"cu1(pi/{1 << (j - i)}) q[{i}],q[{j}];")
    return _emit(n, out)


def bv(n: int, secret: int | None = None) -> str:
    """Bernstein-Vazirani with the all-ones secret by default (matches tests/input/bv_28.qasm)."""
    if secret is None:
        secret = (1 << (n - 1)) - 1
    out = [f"h q[{i}];" for i in range(n - 1)]
    out += [f"x q[{n - 1}];", f"h q[{n - 1}];"]
    out += [f"cx q[{i}],q[{n - 1}];" for i in range(n - 1) if secret >> i & 1]
    out += [f"h q[{i}];" for i in range(n - 1)]
    return _emit(n, out)


HIDDEN_SHIFT_28 = 155182406  # the shift encoded in the reference's tests/output/hidden_shift_28.log


def hidden_shift(n: int, shift: int | None = None, seed: int = 2021) -> str:
    """H^n X^s CZ(2i,2i+1) X^s H^n CZ(2i,2i+1) H^n : |0> -> |s>."""
    assert n % 2 == 0
    if shift is None:
        shift = HIDDEN_SHIFT_28 if n == 28 else random.Random(n * 10000 + seed).getrandbits(n)
    hs = [f"h q[{i}];" for i in range(n)]
    xs = [f"x q[{i}];" for i in range(n) if shift >> i & 1]
    cz = [f"cz q[{2 * i}],q[{2 * i + 1}];" for i in range(n // 2)]
    return _emit(n, hs + xs + cz + xs + hs + cz + hs)


def _grid(n: int):
    rows = max(1, int(math.isqrt(n)))
    while rows > 1 and n % rows and rows * math.ceil(n / rows) - n > rows:
        rows -= 1
    cols = math.ceil(n / rows)
    return rows, cols


def supremacy(n: int, cycles: int = 20, seed: int = 2021) -> str:
    """Google-style random circuit on a near-square grid: h layer, then per cycle one of 8 CZ edge
    patterns followed by a random {rx(pi/2), ry(pi/2), t} on every qubit the CZ layer did not touch
    (never repeating the previous one-qubit gate on that qubit)."""
    rng = random.Random(n * 10000 + seed)
    rows, cols = _grid(n)
    qid = lambda r, c: r * cols + c
    patterns = []
    for kind in range(8):
        edges = []
        for r in range(rows):
            for c in range(cols):
                if kind < 4:      # horizontal edges, class = (c%2, r%2)
                    if c + 1 < cols and (c % 2, r % 2) == (kind & 1, kind >> 1):
                        a, b = qid(r, c), qid(r, c + 1)
                    else:
                        continue
                else:             # vertical edges, class = (r%2, c%2)
                    k = kind - 4
                    if r + 1 < rows and (r % 2, c % 2) == (k & 1, k >> 1):
                        a, b = qid(r, c), qid(r + 1, c)
                    else:
                        continue
                if a < n and b < n:
                    edges.append((a, b))
        patterns.append(edges)
    order = [0, 5, 2, 7, 1, 4, 3, 6]
    out = [f"h q[{i}];" for i in range(n)]
    names = ["rx(pi*0.5)", "ry(pi*0.5)", "t"]
    last = [-1] * n
    for cyc in range(cycles):
        edges = patterns[order[cyc % 8]]
        touched = set()
        for a, b in edges:
            out.append(f"cz q[{a}],q[{b}];")
            touched.update((a, b))
        for q in range(n):
            if q in touched:
                continue
            k = rng.choice([i for i in range(3) if i != last[q]])
            last[q] = k
            out.append(f"{names[k]} q[{q}];")
    return _emit(n, out)


def quantum_volume(n: int, depth: int = 16, seed: int = 2021) -> str:
    """depth layers of random pairings; each pair gets a generic two-qubit block u3 u3 (cx u3 u3) x3."""
    rng = random.Random(n * 10000 + seed)
    ang = lambda: "%.16f" % rng.uniform(0.0, 2 * math.pi)
    out = []
    for _ in range(depth):
        perm = list(range(n))
        rng.shuffle(perm)
        for k in range(n // 2):
            a, b = perm[2 * k], perm[2 * k + 1]
            out.append(f"u3({ang()},{ang()},{ang()}) q[{a}];")
            out.append(f"u3({ang()},{ang()},{ang()}) q[{b}];")
            for _rep in range(3):
                out.append(f"cx q[{a}],q[{b}];")
                out.append(f"u3({ang()},{ang()},{ang()}) q[{a}];")
                out.append(f"u3({ang()},{ang()},{ang()}) q[{b}];")
    return _emit(n, out)


def _regular3(n: int, rng: random.Random):
    assert n % 2 == 0 and n >= 4
    while True:
        stubs = [v for v in range(n) for _ in range(3)]
        rng.shuffle(stubs)
        edges = set()
        ok = True
        for i in range(0, len(stubs), 2):
            a, b = stubs[i], stubs[i + 1]
            if a == b or (min(a, b), max(a, b)) in edges:
                ok = False
                break
            edges.add((min(a, b), max(a, b)))
        if ok:
            return sorted(edges)


def qaoa(n: int, p: int = 2, seed: int = 2021) -> str:
    """MaxCut QAOA on a seeded random 3-regular graph: h all; per level cx-rz-cx per edge, rx on all."""
    rng = random.Random(n * 10000 + seed)
    edges = _regular3(n, rng)
    out = [f"h q[{i}];" for i in range(n)]
    for _ in range(p):
        gamma = "%.16f" % rng.uniform(0.0, 2 * math.pi)
        beta = "%.16f" % rng.uniform(0.0, 2 * math.pi)
        for a, b in edges:
            out += [f"cx q[{a}],q[{b}];", f"rz({gamma}) q[{b}];", f"cx q[{a}],q[{b}];"]
        out += [f"rx({beta}) q[{i}];" for i in range(n)]
    return _emit(n, out)


def adder(n: int, seed: int = 2021) -> str:
    """Cuccaro ripple-carry adder: carry-in q[0], then interleaved b_i,a_i, carry-out q[n-1]; n = 2m+2."""
    assert n % 2 == 0 and n >= 4
    m = (n - 2) // 2
    rng = random.Random(n * 10000 + seed)
    cin, cout = 0, n - 1
    b = [1 + 2 * i for i in range(m)]
    a = [2 + 2 * i for i in range(m)]
    out = []
    for qs in (a, b):
        v = rng.getrandbits(m)
        out += [f"x q[{qs[i]}];" for i in range(m) if v >> i & 1]

    def maj(c, y, x):
        return [f"cx q[{x}],q[{y}];", f"cx q[{x}],q[{c}];", f"ccx q[{c}],q[{y}],q[{x}];"]

    def uma(c, y, x):
        return [f"ccx q[{c}],q[{y}],q[{x}];", f"cx q[{x}],q[{c}];", f"cx q[{c}],q[{y}];"]

    out += maj(cin, b[0], a[0])
    for i in range(1, m):
        out += maj(a[i - 1], b[i], a[i])
    out.append(f"cx q[{a[m - 1]}],q[{cout}];")
    for i in range(m - 1, 0, -1):
        out += uma(a[i - 1], b[i], a[i])
    out += uma(cin, b[0], a[0])
    return _emit(n, out)


def basis_change(n: int, depth: int | None = None, seed: int = 2021) -> str:
    """Fermionic basis-change style Givens network: x on every other qubit, then brick layers of
    Givens(theta) = cx b,a ; cry(2 theta) a,b ; cx b,a  followed by rz(phi) on b."""
    rng = random.Random(n * 10000 + seed)
    depth = n if depth is None else depth
    out = [f"x q[{i}];" for i in range(0, n, 2)]
    for layer in range(depth):
        for a in range(layer % 2, n - 1, 2):
            b = a + 1
            th = "%.16f" % rng.uniform(0.0, 2 * math.pi)
            ph = "%.16f" % rng.uniform(0.0, 2 * math.pi)
            out += [f"cx q[{b}],q[{a}];", f"cry({th}) q[{a}],q[{b}];", f"cx q[{b}],q[{a}];", f"rz({ph}) q[{b}];"]
    return _emit(n, out)


def random_circuit(n: int, ngates: int = 200, seed: int = 0, names: List[str] | None = None) -> str:
    """Every token the parser knows, uniformly at random (test workload; `tdg` excluded by default because
    the reference's own two backends disagree on it, see oracle/oracle.py)."""
    rng = random.Random(seed)
    one = ["h", "x", "y", "z", "s", "sdg", "t"]
    onep = ["rx", "ry", "rz", "u1"]
    two = ["cx", "cy", "cz"]
    twop = ["crx", "cry", "crz", "cu1"]
    pool = names or (one + onep + two + twop + ["u3", "ccx"])
    ang = lambda: "%.16f" % rng.uniform(0.0, 2 * math.pi)
    out = []
    for _ in range(ngates):
        g = rng.choice(pool)
        if g in one or g == "tdg":
            out.append(f"{g} q[{rng.randrange(n)}];")
        elif g in onep:
            out.append(f"{g}({ang()}) q[{rng.randrange(n)}];")
        elif g == "u3":
            out.append(f"u3({ang()},{ang()},{ang()}) q[{rng.randrange(n)}];")
        elif g in two:
            a, b = rng.sample(range(n), 2)
            out.append(f"{g} q[{a}],q[{b}];")
        elif g in twop:
            a, b = rng.sample(range(n), 2)
            out.append(f"{g}({ang()}) q[{a}],q[{b}];")
        elif g == "ccx":
            a, b, c = rng.sample(range(n), 3)
            out.append(f"ccx q[{a}],q[{b}],q[{c}];")
        else:
            raise ValueError(g)
    return _emit(n, out)


FAMILIES: Dict[str, Callable[..., str]] = {
    "qft": qft, "bv": bv, "hidden_shift": hidden_shift, "supremacy": supremacy,
    "quantum_volume": quantum_volume, "qaoa": qaoa, "adder": adder, "basis_change": basis_change,
}


def generate(name: str) -> str:
    """`family_N` -> QASM text, e.g. generate('supremacy_30')."""
    fam, _, n = name.rpartition("_")
    return FAMILIES[fam](int(n))


if __name__ == "__main__":
    import sys
    sys.stdout.write(generate(sys.argv[1]))


def inverse(text: str) -> str:
    """U -> U^dagger in the same dialect (gate order reversed, each gate inverted).  Used for the
    size-independent round-trip property U^dagger U |0> = |0> at full benchmark sizes."""
    lines = [l for l in text.split("\n") if l.strip()]
    head = [l for l in lines if l.split()[0] in ("OPENQASM", "include", "qreg", "//")]
    body = [l for l in lines if l not in head]
    self_inv = {"h", "x", "y", "z", "cx", "cy", "cz", "ccx"}
    swap = {"s": "sdg", "sdg": "s", "t": "tdg", "tdg": "t"}
    out = []
    for l in reversed(body):
        tok, operand = l.split()[0], l.split()[1]
        name = tok.split("(")[0]
        if name in self_inv:
            out.append(l)
        elif name in swap:
            out.append(f"{swap[name]} {operand}")
        else:
            ps = tok[tok.index("(") + 1: tok.rindex(")")].split(",")

            def neg(p):
                if p.startswith("pi"):
                    v = math.pi * float(p[3:]) if p[2] == "*" else math.pi / float(p[3:])
                    return "%.17g" % (-v)
                return p[1:] if p.startswith("-") else "-" + p
            if name == "u3":
                ps = [neg(ps[0]), neg(ps[2]), neg(ps[1])]
            else:
                ps = [neg(p) for p in ps]
            out.append(f"{name}({','.join(ps)}) {operand}")
    return "".join(l + "\n" for l in head + out)
