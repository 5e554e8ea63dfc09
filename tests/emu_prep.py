"""CPU emulation of the K9 (data preparation) part of the kernel namespace `tradeaid_b200.lib` — TEST INFRASTRUCTURE ONLY.

Same Python signatures as lib.prep_*, on CPU tensors. The element arithmetic is the kernels' own (csrc/dataprep_math.cuh
built for the host: tests/host_harness/dataprep_host.cpp), the data movement around it (status words, order statistics,
the open-addressing vocabulary table with its capacity / load-factor contract) is restated with numpy. It lets the CPU
suite run the host logic of tradeaid_b200.data_utils — retry of the vocabulary table, numpy's percentile interpolation,
status -> exception mapping, the chain and the train/val split — without a GPU. The product never imports this module."""
import ctypes
import math
import os
import subprocess
import tempfile

import numpy as np
import torch

PREP_STATUS_WORDS = 4
PREP_NONFINITE, PREP_ZERO_PREV, PREP_OVERFLOW, PREP_TABLE_FULL = 0, 1, 2, 3
PREP_NO_ROW = 2 ** 63 - 1
PREP_MAX_RANKS = 4
PREP_MAX_GROUPS = 8192
PREP_DECADES = 633
PREP_DECADE_MIN = -324
PREP_MAX_DECIMALS = 22
VOCAB_MAX_PROBES = 4096
_EMPTY = np.uint64(0xFFFFFFFFFFFFFFFF)

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_host = None
P = ctypes.c_void_p
stats = {"vocab_inserts": 0}


def host():
    global _host
    if _host is None:
        out = os.path.join(tempfile.mkdtemp(prefix="emu_prep_"), "libdataprep_host.so")
        src = os.path.join(_ROOT, "tests", "host_harness", "dataprep_host.cpp")
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, src], check=True)
        _host = ctypes.CDLL(out)
    return _host


def _np(t):
    assert t.device.type == "cpu" and t.is_contiguous()
    return t.numpy()


def _note(status, word, rows):
    if len(rows):
        status[word] = min(int(status[word]), int(rows[0]))


def prep_status_reset(status):
    status[:PREP_STATUS_WORDS] = PREP_NO_ROW


def prep_percent_changes(x, out, file_starts, decimals, status):
    assert x.data_ptr() != out.data_ptr() and 0 <= decimals <= PREP_MAX_DECIMALS
    xa, oa = _np(x), _np(out)
    _note(status, PREP_NONFINITE, np.flatnonzero(~np.isfinite(xa)))
    starts = file_starts.tolist() + [xa.size]
    zero_rows = []
    for a, b in zip(starts[:-1], starts[1:]):
        seg = np.ascontiguousarray(xa[a:b])
        res = np.empty_like(seg)
        host().h_percent(P(seg.ctypes.data), ctypes.c_longlong(seg.size), ctypes.c_int(decimals), P(res.ctypes.data))
        oa[a:b] = res                                   # h_percent keeps 0.0 after a zero previous value, like the kernel
        zero_rows.extend((a + np.flatnonzero(seg[:-1] == 0)).tolist())
    _note(status, PREP_ZERO_PREV, sorted(zero_rows))


def prep_range(x, out, has_whole_digits, decimals, lower, upper, upper_clamp, thresholds, factors, status):
    xa = np.ascontiguousarray(_np(x))
    bad = np.flatnonzero(~np.isfinite(xa))
    _note(status, PREP_NONFINITE, bad)
    safe = xa.copy()
    safe[bad] = 1.0
    res = np.empty_like(safe)
    thr = _np(thresholds) if has_whole_digits else np.zeros(PREP_DECADES)
    fac = _np(factors) if has_whole_digits else np.ones(PREP_DECADES)
    host().h_range(P(safe.ctypes.data), ctypes.c_longlong(safe.size), ctypes.c_int(1 if has_whole_digits else 0),
                   ctypes.c_int(decimals), ctypes.c_double(lower), ctypes.c_double(upper), ctypes.c_double(upper_clamp),
                   P(thr.ctypes.data), P(fac.ctypes.data), P(res.ctypes.data))
    res[bad] = xa[bad]
    _note(status, PREP_OVERFLOW, np.flatnonzero(~np.isfinite(res) & np.isfinite(xa)))
    _np(out)[:] = res


def prep_range_auto(x, is_int, out, has_whole_digits, lower, upper, upper_clamps, thresholds, factors, status, n_flagged):
    """csrc/dataprep.cu prep_range_auto_kernel with the kernels' own element arithmetic built for the host."""
    xa = np.ascontiguousarray(_np(x))
    bad = np.flatnonzero(~np.isfinite(xa))
    _note(status, PREP_NONFINITE, bad)
    safe = xa.copy()
    safe[bad] = 1.0
    res = np.empty_like(safe)
    flags = np.zeros(safe.size, dtype=np.int32)
    thr = _np(thresholds) if has_whole_digits else np.zeros(PREP_DECADES)
    fac = _np(factors) if has_whole_digits else np.ones(PREP_DECADES)
    clamps = np.asarray([float(v) for v in upper_clamps], dtype=np.float64)
    ints = np.ascontiguousarray(_np(is_int)) if is_int is not None else None
    host().h_range_auto(P(safe.ctypes.data), P(ints.ctypes.data) if ints is not None else None, ctypes.c_longlong(safe.size),
                        ctypes.c_int(1 if has_whole_digits else 0), ctypes.c_double(lower), ctypes.c_double(upper),
                        P(clamps.ctypes.data), P(thr.ctypes.data), P(fac.ctypes.data), P(res.ctypes.data), P(flags.ctypes.data))
    res[bad] = xa[bad]
    flags[bad] = 0
    _note(status, PREP_OVERFLOW, np.flatnonzero(~np.isfinite(res) & np.isfinite(xa) & (flags == 0)))
    _np(out)[:] = res
    n_flagged[0] += int(flags.sum())


def prep_select_workspace_bytes():
    return (2 * PREP_MAX_RANKS + PREP_MAX_RANKS * 256) * 8


def prep_order_statistics(x, ranks, out, workspace, status):
    xa = _np(x)
    _note(status, PREP_NONFINITE, np.flatnonzero(~np.isfinite(xa)))
    srt = np.sort(xa)
    for j, r in enumerate(ranks.tolist()):
        out[j] = float(srt[r])


def prep_minmax_workspace_bytes(N):
    return 24 * max(1, min((N + 511) // 512, 148 * 8))


def prep_filtered_minmax(x, lo, hi, out3, workspace):
    xa = _np(x)
    kept = xa[(xa >= lo) & (xa <= hi)]
    out3[0] = float(kept.min()) if kept.size else math.inf
    out3[1] = float(kept.max()) if kept.size else -math.inf
    out3[2] = float(kept.size)


def prep_bin_assign(x, out, pos_bounds, neg_bounds, bounds_sorted=False):
    xa = np.ascontiguousarray(_np(x))
    G = pos_bounds.numel() - 1
    assert 1 <= G <= PREP_MAX_GROUPS
    res = np.empty(xa.size, dtype=np.int32)
    pos, neg = _np(pos_bounds), _np(neg_bounds)
    host().h_bin(P(xa.ctypes.data), ctypes.c_longlong(xa.size), P(pos.ctypes.data), P(neg.ctypes.data), ctypes.c_int(G),
                 ctypes.c_int(1 if bounds_sorted else 0), P(res.ctypes.data))
    _np(out)[:] = res.astype(np.float64)


def _keys(xa):
    k = np.ascontiguousarray(xa, dtype=np.float64).view(np.uint64).copy()
    k[xa == 0] = 0                                       # 0.0 and -0.0 are one vocabulary entry
    return k


def prep_vocab_insert(x, table, counters, status):
    """Contract of ta_prep_vocab_insert: clears table and counters; counters[0] = distinct values placed; a row whose key
    finds no slot within the probe limit sets status[TABLE_FULL] (the emulation declares the table full when the distinct
    values do not fit, which is when the device's bounded linear probing fails)."""
    stats["vocab_inserts"] += 1
    cap = table.numel()
    assert cap >= 2 and cap & (cap - 1) == 0
    xa = _np(x)
    _note(status, PREP_NONFINITE, np.flatnonzero(~np.isfinite(xa)))
    tab = _np(table).view(np.uint64)
    tab[:] = _EMPTY
    counters[:2] = 0
    keys = _keys(xa[np.isfinite(xa)])
    uniq, first = np.unique(keys, return_index=True)
    if uniq.size > cap:
        placed = uniq[np.argsort(first)][:cap]
        _note(status, PREP_TABLE_FULL, [int(np.sort(first)[cap])])
    else:
        placed = uniq
    tab[:placed.size] = placed                          # slot order is an implementation detail of the device table
    counters[0] = int(placed.size)


def prep_vocab_build(table, n_distinct, keys_ws, slots_ws, vocabulary, token_of_slot, counters):
    tab = _np(table).view(np.uint64)
    occ = np.flatnonzero(tab != _EMPTY)
    assert occ.size == n_distinct
    vals = tab[occ].view(np.float64)
    order = np.argsort(vals, kind="stable")
    _np(vocabulary)[:n_distinct] = vals[order]
    tos = _np(token_of_slot)
    tos[occ[order]] = np.arange(n_distinct, dtype=np.int32)
    counters[1] = n_distinct


def prep_vocab_lookup(x, table, token_of_slot, tokens):
    tab = _np(table).view(np.uint64)
    occ = np.flatnonzero(tab != _EMPTY)
    slot_of = {int(k): int(s) for k, s in zip(tab[occ], occ)}
    tos = _np(token_of_slot)
    keys = _keys(_np(x))
    _np(tokens)[:] = np.array([tos[slot_of[int(k)]] if int(k) in slot_of else -1 for k in keys], dtype=np.int32)
