"""Offline soak of the kernels' exact round (csrc/dataprep_math.cuh py_round, built for the host) against CPython round():
18 M values per seed over ndigits 0..12 (uniform, exact binary ties, decimal half-way cases, 24 decades). Build the harness
first:  g++ -O2 -ffp-contract=off -shared -fPIC -o /tmp/t/libdph.so tests/host_harness/dataprep_host.cpp
usage: tests/soak/dataprep_round_soak.py <seed>      (round 1: seeds 1-3, 5.4e7 values, 0 mismatches)"""
import ctypes, random, struct, numpy as np, sys
L = ctypes.CDLL("/tmp/t/libdph.so"); P = ctypes.c_void_p
rng = np.random.default_rng(int(sys.argv[1]))
bad = tot = 0
for nd in [0, 1, 2, 3, 4, 5, 6, 8, 12]:
    n = 2_000_000
    parts = [rng.uniform(-1000, 1000, n // 4),
             rng.integers(-10**9, 10**9, n // 4) / 2.0 ** rng.integers(0, 34, n // 4),                 # exact binary fractions: ties
             (rng.integers(-10**7, 10**7, n // 4) + 0.5) / 10.0 ** rng.integers(0, 9, n // 4),           # decimal half-way (inexact)
             rng.standard_normal(n // 4) * 10.0 ** rng.integers(-12, 13, n // 4)]
    x = np.ascontiguousarray(np.concatenate(parts)); o = np.empty_like(x)
    L.h_py_round(P(x.ctypes.data), ctypes.c_longlong(x.size), ctypes.c_int(nd), P(o.ctypes.data))
    want = np.array([round(v, nd) for v in x.tolist()])
    bad += int((want.view(np.uint64) != o.view(np.uint64)).sum()); tot += x.size
print("checked", tot, "mismatches", bad)
