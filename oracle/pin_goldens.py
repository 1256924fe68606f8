"""Pin oracle.c/oracle.py against the reference's own golden outputs (n=28, takes minutes on 8 cores).

    python oracle/pin_goldens.py            # prints one PASS/FAIL line per golden, exit 1 on any FAIL

tests/golden/*.log are byte-exact reconstructions of /root/reference/tests/output/*.log (sha256 == LFS oid,
tests/golden/lfs_oids.json); hidden_shift_28.qasm is regenerated (its original is lost) with the shift the
golden output encodes.  TEST INFRASTRUCTURE ONLY.
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:] = [p for p in sys.path if os.path.abspath(p or '.') != os.path.dirname(os.path.abspath(__file__))]
sys.path.insert(0, ROOT)
from oracle import oracle as O
from hyquas_b200 import circuits as C

def main():
    G = os.path.join(ROOT, "tests", "golden")
    cases = {
        "qft_28": open(os.path.join(G, "qft_28.qasm")).read(),
        "bv_28": open(os.path.join(G, "bv_28.qasm")).read(),
        "hidden_shift_28": C.hidden_shift(28),
    }
    bad = 0
    for name, text in cases.items():
        t0 = time.time()
        n, st = O.simulate_qasm(text)
        got = O.dump_state(st, n)
        want = open(os.path.join(G, name + ".log")).read()
        ok = got == want
        bad += not ok
        print("%s %s  (%.1f s, byte-exact=%s)" % ("PASS" if ok else "FAIL", name, time.time() - t0, ok), flush=True)
        del st
    sys.exit(1 if bad else 0)

if __name__ == "__main__":
    main()
