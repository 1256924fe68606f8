#!/usr/bin/env python
"""Phase-by-phase wall clock of the public path (QASM text -> parse -> compile -> run -> dump -> close) for supremacy_30, four
times in one process: shows what the state cache and the small-buffer pool buy (HQ_STATE_CACHE=0 to compare; profiles/
r01_s22_e2e_probe_nocache.log vs r01_s23_e2e_probe.log)."""
import time, sys, os
sys.path.insert(0, "/root/repo")
from hyquas_b200 import api, circuits as C
api.init()
text = C.generate("supremacy_30")
for it in range(4):
    t0=time.perf_counter(); ce = api.Circuit.from_qasm(text); t1=time.perf_counter(); ce.compile(); t2=time.perf_counter()
    ce.run(copy_back=False, destroy=False); t3=time.perf_counter(); d=ce.dump(); t4=time.perf_counter(); io=ce.io_bytes(); t5=time.perf_counter(); ce.close(); t6=time.perf_counter()
    print("iter",it,"parse %.1f compile %.1f run %.1f dump %.1f io %.1f close %.1f ms"%tuple(1e3*x for x in (t1-t0,t2-t1,t3-t2,t4-t3,t5-t4,t6-t5)), flush=True)
