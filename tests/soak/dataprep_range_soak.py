"""Offline soak of the kernels' ranging arithmetic (range_element, built for the host) against the oracle restatement of
range_numeric_data: 2.7 M values per seed over 9 (whole digits, decimals) settings. Same harness as dataprep_round_soak.py.
usage: tests/soak/dataprep_range_soak.py <seed>      (round 1: seeds 5-6, 5.4e6 values, 0 mismatches)"""
import ctypes, sys, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import dataprep as O
from tradeaid_b200 import data_utils as du
L = ctypes.CDLL("/tmp/t/libdph.so"); P = ctypes.c_void_p
rng = np.random.default_rng(int(sys.argv[1]))
thr = np.array(du._log10_thresholds())
bad = tot = 0
for nwd, dp in [(1, 2), (2, 1), (3, 0), (None, 2), (0, 3), (-2, 5), (4, 4), (2, 0), (1, 6)]:
    n = 300_000
    x = np.concatenate([rng.uniform(-500, 500, n // 3), rng.standard_normal(n // 3) * 10.0 ** rng.integers(-8, 9, n // 3),
                        np.round(rng.uniform(-100, 100, n // 3), 2)])
    x[::1000] = 0.0
    x = np.ascontiguousarray(x); o = np.empty_like(x)
    fac = np.array(du._decade_factors(nwd if nwd is not None else 1))
    lo = up = cl = 0.0
    if nwd is not None:
        lo, up = float(10 ** (nwd - 1)), float(10 ** nwd)
        cl = float(10 ** nwd - (10 ** (-dp))) if dp > 0 else float(10 ** nwd - 1)
    L.h_range(P(x.ctypes.data), ctypes.c_longlong(x.size), ctypes.c_int(nwd is not None), ctypes.c_int(dp), ctypes.c_double(lo),
              ctypes.c_double(up), ctypes.c_double(cl), P(thr.ctypes.data), P(fac.ctypes.data), P(o.ctypes.data))
    want = np.array([float(v) for v in O.range_values(x.tolist(), nwd, dp)])
    bad += int((want.view(np.uint64) != o.view(np.uint64)).sum()); tot += x.size
print("range checked", tot, "mismatches", bad)
