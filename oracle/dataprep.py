"""oracle/dataprep.py — CPU restatement of the reference's one-shot data preparation.
TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing in the product package); it is the checker for the K9 kernels.

Restates, in plain Python floats (the reference's own arithmetic: CPython round(), math.log10, int/float powers) and
numpy's percentile, the algorithm of

  /root/reference/data_utils.py:612-662   convert_to_percent_changes   (file_cache.py:329-385: same, zero -> 0.0)
  /root/reference/data_utils.py:361-470   range_numeric_data
  /root/reference/data_utils.py:473-609   bin_numeric_data
  /root/reference/data_utils.py:212-225   numerical_representation

Pinning: the reference's tests hold no vectors for these functions; tests/golden/make_golden_dataprep.py imports the
reference's data_utils in the build container, runs it on seeded columns and stores inputs + outputs in
tests/golden/dataprep.pt; tests/test_dataprep_cpu.py holds this file to those fixtures bit for bit.
Pure-Python loops: meant for columns of a few thousand rows.
"""
from __future__ import annotations

import math

import numpy as np


def percent_changes(data, decimals=2, on_zero="raise"):
    """Backward-looking change in percent, first row 0.0 (data_utils.py:645-657). on_zero='keep' is the
    file_cache.py:366-372 behaviour (0.0 and carry on) instead of the ZeroDivisionError of data_utils.py:652."""
    decimals = 2 if decimals is None else decimals
    out = [0.0]
    for prev, cur in zip(data[:-1], data[1:]):
        if prev == 0:
            if on_zero == "raise":
                raise ZeroDivisionError(f"Cannot calculate percentage change: previous value is zero at index {len(out) - 1}.")
            out.append(0.0)
            continue
        out.append(round(((cur - prev) / prev) * 100, decimals))
    return out


def percent_changes_by_file(data, file_lengths, decimals=2, on_zero="raise"):
    """The loader converts every file on its own (data_utils.py:148) and concatenates."""
    out, at = [], 0
    for n in file_lengths:
        out.extend(percent_changes(data[at:at + n], decimals, on_zero))
        at += n
    return out


def range_values(data, whole_digits, decimals):
    """Scale every value into `whole_digits` digits before the point and round to `decimals` (data_utils.py:421-457).
    decimals=None: every element is rounded to the length of the fraction part of its own str() (:431)."""
    out = []
    fixed = decimals
    for v in data:
        decade = 0 if v == 0 else int(math.floor(math.log10(abs(v))))          # :423-429
        decimals = fixed if fixed is not None else (len(str(v).split('.')[-1]) if '.' in str(v) else 0)   # :431
        decimals = max(0, decimals)
        factor = 1 if whole_digits is None else 10 ** ((whole_digits - 1) - decade)   # :436-440
        s = round(v * factor, decimals) if factor != 0 else 0.0                # :443
        if whole_digits is not None:                                           # :445-455
            a = abs(s)
            if 0 < a < 10 ** (whole_digits - 1):
                a = 10 ** (whole_digits - 1)
            if a >= 10 ** whole_digits:
                a = 10 ** whole_digits - (10 ** (-decimals)) if decimals > 0 else 10 ** whole_digits - 1
            s = a * (-1 if v < 0 else 1)
        out.append(s)
    return out


def group_boundaries(num_groups, exponent, max_abs):
    pos = [0.0] + [((i / num_groups) ** exponent) * max_abs for i in range(1, num_groups + 1)]   # :514-521
    neg = [-b for b in reversed(pos[1:])] + [0.0]                                              # :524
    return pos, neg


def bin_groups(data, num_groups, outlier_percentile=5, exponent=2.0):
    """Signed group numbers -G..G (data_utils.py:503-557)."""
    lo = np.percentile(data, outlier_percentile)
    hi = np.percentile(data, 100 - outlier_percentile)
    kept = [v for v in data if lo <= v <= hi]
    if not kept:
        raise ValueError("All data points were filtered out as outliers.")
    pos, neg = group_boundaries(num_groups, exponent, max(abs(min(kept)), abs(max(kept))))
    out = []
    for v in data:
        if v > 0:
            g = next((j + 1 for j in range(num_groups) if pos[j] <= v < pos[j + 1]), num_groups)
        elif v == 0:
            g = 0
        else:
            g = next((-(num_groups - j) for j in range(num_groups) if neg[j] <= v < neg[j + 1]), -num_groups)
        out.append(g)
    return out


def vocabulary_tokens(data):
    """(tokens, vocabulary) with vocabulary = the distinct values ascending (data_utils.py:222-224)."""
    vocabulary = sorted(set(data))
    index = {v: i for i, v in enumerate(vocabulary)}
    return [index[v] for v in data], vocabulary


def prepare_modality(column, params, file_lengths=None, on_zero="keep"):
    """The per-modality chain of main.py:100-210 + :278 on the legacy parameter list (schema.py:213-248). main.py loads
    through file_cache.load_file_data_cached, whose percent conversion keeps 0.0 after a zero (on_zero='keep')."""
    p = list(params) + [None] * (12 - len(params))
    x = list(column)
    if p[3] is True:
        x = percent_changes_by_file(x, file_lengths or [len(x)], p[5] if p[5] else 2, on_zero)
    if p[4] is not None or p[5] is not None:
        x = range_values(x, p[4], p[5])
    if p[6] is not None:
        x = bin_groups(x, p[6], p[10] if p[10] is not None else 0.1, p[11] if p[11] is not None else 2.2)
    return vocabulary_tokens(x)
