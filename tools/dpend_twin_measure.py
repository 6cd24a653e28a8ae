"""Measure the shadow-twin constant of the chaotic double-pendulum config on the bench's own inputs:
GPU (contract and strict) against the CPU checker on the first 256 instances of W.dpend_ensemble(seed=2), 10 000 steps,
snapshots every 1000 steps; prints max over instances and snapshots of  (err - 1e-9) / twin  where twin = separation of the
checker's own run from its run with theta1 nudged by one ulp.  tests/test_gpu_parity.py uses 10x the printed maximum."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.oracle_py import Checker, build  # noqa: E402
from sopot_b200 import workloads as W  # noqa: E402
from sopot_b200.capi import FP_CONTRACT, FP_STRICT, Engine  # noqa: E402

build()
port = Checker("port")
n = 256
w = W.dpend_ensemble(1 << 22, seed=2)
keys = ("m1", "m2", "L1", "L2", "g", "th1", "th2", "w1", "w2")
args = [np.ascontiguousarray(w[k][:n]) for k in keys]
_, ref, _ = port.dpend_rk4(*args, 0.0, 10.0, 1e-3, store_every=1000, n_snap=10, nthreads=8)
nudged = list(args)
nudged[5] = np.nextafter(args[5], np.inf)
_, twin_t, _ = port.dpend_rk4(*nudged, 0.0, 10.0, 1e-3, store_every=1000, n_snap=10, nthreads=8)


def rel(a, b):
    return np.max(np.abs(a - b), axis=1) / np.maximum(np.max(np.abs(b), axis=1), 1e-300)     # [snap][instance]


twin = rel(twin_t, ref)
eng = Engine(0)
for name, mode in (("contract", FP_CONTRACT), ("strict", FP_STRICT)):
    _, traj, st = eng.dpend_rk4(*args, n, 0.0, 1e-3, 10000, fp_mode=mode, store_every=1000)
    err = rel(traj, ref)
    ratio = np.maximum(err - 1e-9, 0.0) / np.maximum(twin, 1e-300)
    print(name, "max err at t=1s %.3e" % err[0].max(), " max (err-1e-9)/twin %.3f" % ratio.max(),
          " median err/twin at t=10s %.3f" % float(np.median(err[-1] / np.maximum(twin[-1], 1e-300))),
          " max err/twin over all %.3f" % float((err / np.maximum(twin, 1e-300)).max()),
          " instances with err>1e-9 at 10 s: %d" % int((err[-1] > 1e-9).sum()))
eng.close()
