"""Operation count of the engine's OWN rocket derivative by the counting-scalar method.

    python tools/build_variant.py opcount rocket.cu -DSB_COUNT_OPS          (CPU box: builds the instrumented library)
    SOPOT_B200_LIB=sopot_b200/lib/variants/libsopot_b200_opcount.so python tools/rocket_opcount.py > profiles/rocket_opcount.json

In the instrumented build every strict-mode + - * / sqrt and every transcendental call executed by thread 0 of block 0 is
tallied (common.cuh, SB_COUNT_OPS).  Two measurements with the bench's own tables and the nominal launch (elevation 85 deg):
  * one derivative evaluation in powered flight at t = 2 s (the state the reference count of 411 flop was taken at), and
  * the whole 30 s flight, 3000 RK4 steps in the reference association (stage arithmetic and running statistics included):
    the flight-phase-weighted flop per instance-step that bench.py uses as the numerator of the rocket roofline.
Each operation counts 1 flop (the reference's convention; no FMA in strict mode)."""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sopot_b200 import workloads as W  # noqa: E402
from sopot_b200.capi import FP_STRICT, Engine  # noqa: E402

KINDS = ("add_sub", "mul", "div", "sqrt", "transcendental")
eng = Engine(0)
fn = eng.lib.sopot_b200_debug_opcounts
fn.restype, fn.argtypes = ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]


def read(reset=True):
    buf = (ctypes.c_ulonglong * 8)()
    eng._ck(fn(eng.ctx, buf, int(reset)))
    return {k: int(buf[i]) for i, k in enumerate(KINDS)}


w = W.rocket_ensemble(1)
one = (np.array([85.0]), np.array([0.0]), np.array([0.16]), np.array([9.80665]))
read()
# the nominal flight, stored every step, to pick a powered-flight state
st, stats, traj, _ = eng.rocket_rk4(*one, w["engine"], w["mass_tab"], None, None, 1, 0.01, 3000, fp_mode=FP_STRICT, store_every=1)
flight = read()
state_2s = np.ascontiguousarray(traj[199].reshape(14, 1))
eng.rocket_rhs(85.0, 0.0, 0.16, 9.80665, w["engine"], w["mass_tab"], state_2s, fp_mode=FP_STRICT)
rhs_powered = read()
state_coast = np.ascontiguousarray(traj[999].reshape(14, 1))
eng.rocket_rhs(85.0, 0.0, 0.16, 9.80665, w["engine"], w["mass_tab"], state_coast, fp_mode=FP_STRICT)
rhs_coast = read()
total = sum(flight.values())
out = {
    "method": "counting scalar on the device: strict-mode operations of one trajectory (tools/rocket_opcount.py, SB_COUNT_OPS build)",
    "flop_per_instance_step": total / 3000.0,
    "flight_total": flight, "steps": 3000,
    "rhs_powered_flight_t2s": dict(rhs_powered, total=sum(rhs_powered.values())),
    "rhs_coast_t10s": dict(rhs_coast, total=sum(rhs_coast.values())),
    "reference_as_written": {"rhs_powered": 411, "rhs_coast": 342, "per_instance_step": 1870, "source": "SURVEY.md section 8d"},
    "apogee_m": float(stats[0, 0]),
}
print(json.dumps(out, indent=1))
eng.close()
