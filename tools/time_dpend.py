"""Time the double-pendulum ensemble kernel of the library selected by SOPOT_B200_LIB (variant sweeps)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sopot_b200 import workloads as W
from sopot_b200.capi import FP_CONTRACT, MEM_DEVICE, Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
eng = Engine(0)
w = W.dpend_ensemble(n)
keys = ("m1", "m2", "L1", "L2", "g", "th1", "th2", "w1", "w2")
d = [eng.to_device(w[k]) for k in keys]
st = eng.empty((4, n))
best = 1e9
for r in range(4):
    eng.event_record(0)
    eng.dpend_rk4(*d, n, 0.0, 1e-3, steps, mem=MEM_DEVICE, fp_mode=FP_CONTRACT, want_status=False, state_out=st, sync=False)
    eng.event_record(1)
    ms = eng.elapsed_ms(0, 1)
    if r: best = min(best, ms)
h = st.download()
print("%-40s %8.2f ms  %.3e inst-steps/s  checksum %.12e" % (os.path.basename(os.environ.get("SOPOT_B200_LIB", "stock")), best, n * steps / best * 1e3, float(np.sum(h[:, ::4097]))))
eng.close()
