"""One-shot data preparation on the device: the reference's `data_utils` processing functions
with the same names, arguments, validation and error behaviour, computed by the K9 kernels (csrc/dataprep.cu) on a
float64 column resident in HBM.

    convert_to_percent_changes   /root/reference/data_utils.py:612-662  (file_cache.py:329-385 with on_zero="keep")
    range_numeric_data           /root/reference/data_utils.py:361-470
    bin_numeric_data             /root/reference/data_utils.py:473-609
    numerical_representation     /root/reference/data_utils.py:212-225
    create_train_val_datasets    /root/reference/data_utils.py:228-290   (slicing only: host)

Every function takes what the reference takes — a Python list — and then returns what the reference returns (lists), or
a float64 CUDA tensor [N], and then returns device tensors: that is the form `prepare_modality` chains, so that a column
goes from raw values to a row of the feeder's token store without a per-element Python loop or a host copy.
Results are BIT-exact to the reference's Python floats (tests/test_dataprep_*.py): equal values share a vocabulary entry
and the entry's index is the token id, so "close" is not good enough. What the host contributes is scalar work done with
Python's own expressions — decade tables from this machine's `math.log10` and `10**k`, `np.percentile`'s interpolation of
four order statistics, the group boundaries — everything proportional to N runs in the kernels.
There is no CPU fallback: without a CUDA device or the library these functions raise.

Non-numeric (string) modalities are tokenised too: their vocabulary is host work by nature (_symbolic_representation), the
token ids go to the device store like every other modality's.
`range_numeric_data(decimal_places=None)` - every element rounded to the length of the fraction its own `str()` shows
(data_utils.py:431) - runs on the device too (_range_auto). What still raises NotImplementedError are parameter values outside
the double / table range (decimal_places > 22, num_groups > 8192).
"""
from __future__ import annotations

import math
import numbers

import numpy as np
import torch

# ---------------------------------------------------------------------------------------------------- host scalars
_DECADE_MIN, _DECADE_MAX = -324, 308
_thresholds_cache = None


def _prev_double(x):
    return math.nextafter(x, -math.inf)


def _log10_thresholds():
    """thresholds[k] = the smallest positive double t with floor(math.log10(t)) >= k + _DECADE_MIN, found with THIS
    machine's libm — the reference calls math.log10 (data_utils.py:429) and glibc rounds arguments a few ulps below a
    power of ten up to the integer, so the decade boundary is not the power of ten itself."""
    global _thresholds_cache
    if _thresholds_cache is None:
        out = []
        for p in range(_DECADE_MIN, _DECADE_MAX + 1):
            t = float(f"1e{p}") if p > -324 else 5e-324
            while t > 5e-324 and math.floor(math.log10(_prev_double(t))) >= p:
                t = _prev_double(t)
            while math.floor(math.log10(t)) < p:
                t = math.nextafter(t, math.inf)
            out.append(t)
        _thresholds_cache = out
    return _thresholds_cache


def _decade_factors(num_whole_digits):
    """factors[k] = what `element * 10**(desired_power - power)` multiplies a FLOAT element with (data_utils.py:436-440):
    an int power converted to double for exponents >= 0, Python's float power below; inf where not representable."""
    out = []
    for p in range(_DECADE_MIN, _DECADE_MAX + 1):
        try:
            out.append(float(10 ** ((num_whole_digits - 1) - p)))
        except OverflowError:
            out.append(math.inf)
    return out


def _percentile_plan(n, q):
    """np.percentile(a, q) for 1-d `a` of length n, method 'linear' (numpy/lib/_function_base_impl.py `_quantile`):
    returns (previous_index, next_index, gamma) in terms of the ascending order statistics."""
    quantile = np.true_divide(q, np.float64(100))
    virtual = (n - 1) * quantile
    prev = int(np.floor(virtual))
    nxt = prev + 1
    gamma = virtual - np.floor(virtual)
    if virtual >= n - 1:                    # at or above the last index: numpy takes the maximum for both neighbours
        gamma = virtual - (-1)
        prev = nxt = n - 1
    if virtual < 0:
        prev = nxt = 0
    return prev, nxt, float(gamma)


def _lerp(a, b, t):
    """numpy's `_lerp`: a + (b - a) * t, evaluated from the other end for t >= 0.5."""
    diff = b - a
    return b - diff * (1 - t) if t >= 0.5 else a + diff * t


def _group_boundaries(num_groups, exponent, max_abs_value):
    """data_utils.py:514-524, verbatim arithmetic."""
    pos = [0.0]
    for i in range(1, num_groups + 1):
        normalized_i = i / num_groups
        scaled_position = normalized_i ** exponent
        pos.append(scaled_position * max_abs_value)
    neg = [-b for b in reversed(pos[1:])] + [0.0]
    return pos, neg


# ---------------------------------------------------------------------------------------------------- device plumbing
_ops_override = None   # tests inject a CPU emulation of the kernel namespace (tests/emu_prep.py); the product leaves this None


def _ops():
    if _ops_override is not None:
        return _ops_override
    from . import lib
    return lib


def _device():
    if _ops_override is not None:
        return torch.device("cpu")
    if not torch.cuda.is_available():
        raise RuntimeError("tradeaid_b200.data_utils runs on a CUDA device (sm_100a); there is no CPU path — "
                           "use the reference's data_utils on a machine without one.")
    return torch.device("cuda", torch.cuda.current_device())


def _first_non_number(data):
    for i, item in enumerate(data):
        if not isinstance(item, numbers.Number):
            return i
    return -1


def _upload(data):
    """Python list of numbers -> float64 device column (floats are doubles; ints are exact up to 2^53)."""
    arr = np.asarray(data, dtype=np.float64)
    if arr.ndim != 1:
        raise ValueError("data must be one-dimensional")
    return torch.from_numpy(arr).to(_device())


class _Status:
    """The device status words of a step (first offending row per condition), read back once."""

    def __init__(self, ops, device):
        self.ops = ops
        self.words = torch.empty(ops.PREP_STATUS_WORDS, dtype=torch.int64, device=device)
        ops.prep_status_reset(self.words)

    def read(self):
        return [None if w == self.ops.PREP_NO_ROW else w for w in self.words.tolist()]

    def raise_nonfinite(self, row, x):
        if row is not None:
            raise ValueError(f"cannot process non-finite value {float(x[row])!r} at index {row}")


def _is_column(x):
    return torch.is_tensor(x)


def _check_column(x, name):
    if x.dtype != torch.float64 or x.dim() != 1 or not (x.is_cuda or _ops_override is not None):
        raise TypeError(f"'{name}' as a tensor must be a float64 CUDA tensor of shape [N]")
    if x.numel() == 0:
        raise ValueError(f"'{name}' must be non-empty.")
    return x.contiguous()


# ---------------------------------------------------------------------------------------------------- percent changes
def convert_to_percent_changes(data, decimal_places=2, file_lengths=None, on_zero="raise"):
    """data_utils.convert_to_percent_changes (data_utils.py:612-662). `file_lengths` (extension, default one file): the
    reference converts file by file (data_utils.py:148, file_cache.py:312-320), so the first row of every file is 0.0.
    on_zero="raise": ZeroDivisionError with the reference's message (data_utils.py:652); "keep": file_cache's variant,
    0.0 for that row (file_cache.py:366-372)."""
    as_list = not _is_column(data)
    if as_list:
        if not isinstance(data, list) or not data:
            raise ValueError("'data' must be a non-empty list.")
        i = _first_non_number(data)
        if i >= 0:
            raise ValueError(f"All elements in 'data' must be numeric. Element at index {i} is "
                             f"{type(data[i]).__name__}: '{data[i]}'.")
    if decimal_places is not None:
        if not isinstance(decimal_places, int) or decimal_places < 0:
            raise ValueError("'decimal_places' must be a non-negative integer or null.")
    else:
        decimal_places = 2
    if on_zero not in ("raise", "keep"):
        raise ValueError("on_zero must be 'raise' or 'keep'")
    ops = _ops()
    if decimal_places > ops.PREP_MAX_DECIMALS:
        raise NotImplementedError(f"decimal_places > {ops.PREP_MAX_DECIMALS} is not supported on the device")
    x = _upload(data) if as_list else _check_column(data, "data")
    n = x.numel()
    lengths = [n] if file_lengths is None else [int(v) for v in file_lengths]
    if sum(lengths) != n or any(v <= 0 for v in lengths):
        raise ValueError(f"Sum of file_lengths ({sum(lengths)}) does not match length of data ({n}).")
    starts = torch.tensor(np.concatenate(([0], np.cumsum(lengths[:-1]))).astype(np.int64), device=x.device)
    st = _Status(ops, x.device)
    out = torch.empty_like(x)
    ops.prep_percent_changes(x, out, starts, decimal_places, st.words)
    nonfinite, zero_prev, _, _ = st.read()
    st.raise_nonfinite(nonfinite, x)
    if zero_prev is not None:
        if on_zero == "raise":
            local = zero_prev - max(int(v) for v in starts.tolist() if v <= zero_prev)   # the reference converts file by file
            raise ZeroDivisionError(f"Cannot calculate percentage change: previous value is zero at index {local}.")
        print(f"Warning: Zero value found (first at index {zero_prev}) causes division by zero. "
              f"Using 0.0% change for the rows that follow a zero.")       # file_cache.py:367-370 prints one line per row
    return out.tolist() if as_list else out


# ---------------------------------------------------------------------------------------------------- ranging
def range_numeric_data(numeric_data, num_whole_digits, decimal_places, out=None):
    """data_utils.range_numeric_data (data_utils.py:361-470). Tensor input: `out` may be the input itself (in place)."""
    as_list = not _is_column(numeric_data)
    if as_list:
        if not isinstance(numeric_data, list):
            raise TypeError("'numeric_data' must be a list.")
        if not numeric_data:
            raise TypeError("'numeric_data' must be a non-empty list.")
        i = _first_non_number(numeric_data)
        if i >= 0:
            raise IndexError(f"Element at index {i} in 'numeric_data' is not a number.")
    if num_whole_digits is not None and not isinstance(num_whole_digits, int):
        raise TypeError("'num_whole_digits' must be an integer or None.")
    if decimal_places is not None and not isinstance(decimal_places, int):
        raise TypeError("'decimal_places' must be an integer or None.")
    if decimal_places is not None and decimal_places < 0:
        raise ValueError("'decimal_places' must be greater than or equal to 0.")
    ops = _ops()
    if decimal_places is None:
        return _range_auto(ops, numeric_data, as_list, num_whole_digits, out)
    if decimal_places > ops.PREP_MAX_DECIMALS:
        raise NotImplementedError(f"decimal_places > {ops.PREP_MAX_DECIMALS} is not supported on the device")
    x = _upload(numeric_data) if as_list else _check_column(numeric_data, "numeric_data")
    lower = upper = upper_clamp = 0.0
    thresholds = factors = None
    if num_whole_digits is not None:
        try:
            lower = float(10 ** (num_whole_digits - 1))
            upper = float(10 ** num_whole_digits)
            if decimal_places > 0:
                upper_clamp = float(10 ** num_whole_digits - (10 ** (-decimal_places)))
            else:
                upper_clamp = float(10 ** num_whole_digits - 1)
        except OverflowError as e:
            raise NotImplementedError(f"num_whole_digits={num_whole_digits} is outside the double range") from e
        thresholds = torch.tensor(_log10_thresholds(), dtype=torch.float64, device=x.device)
        factors = torch.tensor(_decade_factors(num_whole_digits), dtype=torch.float64, device=x.device)
    st = _Status(ops, x.device)
    res = torch.empty_like(x) if out is None else out
    ops.prep_range(x, res, num_whole_digits is not None, decimal_places, lower, upper, upper_clamp, thresholds, factors,
                   st.words)
    nonfinite, _, overflow, _ = st.read()
    st.raise_nonfinite(nonfinite, x)
    if overflow is not None:
        raise OverflowError(f"int too large to convert to float (scaling factor of element {overflow})")
    return res.tolist() if as_list else res


def _range_auto(ops, numeric_data, as_list, num_whole_digits, out):
    """range_numeric_data with decimal_places=None (data_utils.py:431): every element is rounded to the number of characters
    its own str() shows after the '.', so 12.5 keeps one decimal and 0.125 three. On the device that number is found as the
    smallest count of decimals at which round() reproduces the value (what repr's shortest digit string amounts to in its
    positional range; a Python int has none). The few values repr() prints in exponent notation come back flagged and are
    finished here with the reference's own expression."""
    x = _upload(numeric_data) if as_list else _check_column(numeric_data, "numeric_data")
    dev = x.device
    is_int = None
    if as_list and any(isinstance(v, numbers.Integral) for v in numeric_data):
        is_int = torch.from_numpy(np.fromiter((isinstance(v, numbers.Integral) for v in numeric_data), dtype=np.uint8,
                                              count=len(numeric_data))).to(dev)
    nd = ops.PREP_MAX_DECIMALS
    lower = upper = 0.0
    clamps = [0.0] * (nd + 1)
    thresholds = factors = None
    if num_whole_digits is not None:
        try:
            lower = float(10 ** (num_whole_digits - 1))
            upper = float(10 ** num_whole_digits)
            clamps = [float(10 ** num_whole_digits - (10 ** (-d))) if d > 0 else float(10 ** num_whole_digits - 1)
                      for d in range(nd + 1)]
        except OverflowError as e:
            raise NotImplementedError(f"num_whole_digits={num_whole_digits} is outside the double range") from e
        thresholds = torch.tensor(_log10_thresholds(), dtype=torch.float64, device=dev)
        factors = torch.tensor(_decade_factors(num_whole_digits), dtype=torch.float64, device=dev)
    st = _Status(ops, dev)
    flagged = torch.zeros(1, dtype=torch.int64, device=dev)
    res = torch.empty_like(x) if out is None else out
    keep = ops.prep_range_auto(x, is_int, res, num_whole_digits is not None, lower, upper, clamps, thresholds, factors,
                               st.words, flagged)
    nonfinite, _, overflow, _ = st.read()
    del keep
    st.raise_nonfinite(nonfinite, x)
    if overflow is not None:
        raise OverflowError(f"int too large to convert to float (scaling factor of element {overflow})")
    if int(flagged[0]) > 0:
        # magnitudes repr() prints as 1.5e-05 / 1e+16: "5e-05" counts as five "decimals" in the reference. Python's own code.
        rows = torch.nonzero(torch.isnan(res) & ~torch.isnan(x)).flatten().tolist()
        src = numeric_data if as_list else None
        for i in rows:
            v = src[i] if src is not None else float(x[i])
            res[i] = _range_one_by_str(v, num_whole_digits)
    return res.tolist() if as_list else res


def _range_one_by_str(element, num_whole_digits):
    """One element of range_numeric_data(decimal_places=None), the reference's expressions (data_utils.py:421-457)."""
    power_of_10 = 0 if element == 0 else int(math.floor(math.log10(abs(element))))
    apply_dec_places = max(0, len(str(element).split('.')[-1]) if '.' in str(element) else 0)
    scaling_factor = 1
    if num_whole_digits is not None:
        scaling_factor = 10 ** ((num_whole_digits - 1) - power_of_10)
    scaled_value = round(element * scaling_factor, apply_dec_places) if scaling_factor != 0 else 0.0
    if num_whole_digits is not None:
        abs_scaled_value = abs(scaled_value)
        if abs_scaled_value < 10 ** (num_whole_digits - 1) and abs_scaled_value > 0:
            abs_scaled_value = 10 ** (num_whole_digits - 1)
        if abs_scaled_value >= 10 ** num_whole_digits:
            abs_scaled_value = 10 ** num_whole_digits - (10 ** (-apply_dec_places)) if apply_dec_places > 0 else \
                10 ** num_whole_digits - 1
        scaled_value = abs_scaled_value * (-1 if element < 0 else 1)
    return float(scaled_value)


# ---------------------------------------------------------------------------------------------------- binning
def _order_statistics(ops, x, ranks, st):
    r = torch.tensor(ranks, dtype=torch.int64, device=x.device)
    out = torch.empty(len(ranks), dtype=torch.float64, device=x.device)
    ws = torch.empty(ops.prep_select_workspace_bytes(), dtype=torch.uint8, device=x.device)
    ops.prep_order_statistics(x, r, out, ws, st.words)
    return out.tolist()


def percentiles(x, qs):
    """np.percentile(x, q) for each q (linear interpolation) of a float64 device column: the neighbouring order
    statistics come from one radix-select pass set, the interpolation is numpy's."""
    ops = _ops()
    x = _check_column(x, "x")
    st = _Status(ops, x.device)
    n = x.numel()
    plans = [_percentile_plan(n, q) for q in qs]
    out = []
    for k in range(0, len(plans), 2):                      # two percentiles (four ranks) per select
        chunk = plans[k:k + 2]
        ranks = [i for p in chunk for i in (p[0], p[1])]
        vals = _order_statistics(ops, x, ranks, st)
        for j, p in enumerate(chunk):
            out.append(_lerp(vals[2 * j], vals[2 * j + 1], p[2]))
    nonfinite = st.read()[0]
    st.raise_nonfinite(nonfinite, x)
    return out


def bin_numeric_data(data, num_groups, outlier_percentile=5, exponent=2.0, out=None, verbose=True, as_float=False):
    """data_utils.bin_numeric_data (data_utils.py:473-609): groups -num_groups..num_groups with power-law boundaries
    up to the largest magnitude inside the [p, 100-p] percentile band. List input returns a list of ints like the
    reference; tensor input returns int64 groups on the device (float64 with as_float=True: the form
    numerical_representation takes). Prints the reference's breakdown unless verbose=False."""
    as_list = not _is_column(data)
    if as_list:
        if not isinstance(data, list) or not data:
            raise ValueError("'data' must be a non-empty list.")
        i = _first_non_number(data)
        if i >= 0:
            raise ValueError(f"All elements in 'data' must be numeric. Element at index {i} is "
                             f"{type(data[i]).__name__}: '{data[i]}'.")
    if not isinstance(num_groups, int) or num_groups <= 0:
        raise ValueError("'num_groups' must be a positive integer.")
    if not isinstance(outlier_percentile, (int, float)) or not (0 <= outlier_percentile <= 50):
        raise ValueError("'outlier_percentile' must be a number between 0 and 50.")
    if not isinstance(exponent, (int, float)) or exponent < 1:
        raise ValueError("'exponent' must be a number >= 1.")
    ops = _ops()
    if num_groups > ops.PREP_MAX_GROUPS:
        raise NotImplementedError(f"num_groups > {ops.PREP_MAX_GROUPS} is not supported on the device")
    x = _upload(data) if as_list else _check_column(data, "data")
    lower_percentile, upper_percentile = percentiles(x, [outlier_percentile, 100 - outlier_percentile])
    mm = torch.empty(3, dtype=torch.float64, device=x.device)
    ws = torch.empty(ops.prep_minmax_workspace_bytes(x.numel()), dtype=torch.uint8, device=x.device)
    ops.prep_filtered_minmax(x, lower_percentile, upper_percentile, mm, ws)
    fmin, fmax, count = mm.tolist()
    if count == 0:
        raise ValueError("All data points were filtered out as outliers.")
    max_abs_value = max(abs(fmin), abs(fmax))
    pos, neg = _group_boundaries(num_groups, exponent, max_abs_value)
    res = torch.empty_like(x) if out is None else out
    monotone = all(a <= b for a, b in zip(pos, pos[1:])) and all(a <= b for a, b in zip(neg, neg[1:]))
    ops.prep_bin_assign(x, res, torch.tensor(pos, dtype=torch.float64, device=x.device),
                        torch.tensor(neg, dtype=torch.float64, device=x.device), bounds_sorted=monotone)
    if verbose:
        counts = torch.bincount((res + num_groups).to(torch.int64), minlength=2 * num_groups + 1).tolist()
        _print_bin_breakdown(counts, num_groups, pos, neg, x.numel())
    if as_list:
        return [int(v) for v in res.tolist()]
    return res if as_float else res.to(torch.int64)


def _print_bin_breakdown(counts, num_groups, pos, neg, n):
    """The report of data_utils.py:566-607 (same text), from the per-group counts."""
    print("    -> Binning breakdown (only populated bins showing):")
    for i in range(-num_groups, 0):
        c = counts[i + num_groups]
        if c:
            j = num_groups + i
            if i == -num_groups:
                print(f"      Bin {i}: (-inf, {neg[j + 1]:.3f}) - {c} elements")
            else:
                print(f"      Bin {i}: [{neg[j]:.3f}, {neg[j + 1]:.3f}) - {c} elements")
    if counts[num_groups]:
        print(f"      Bin  0: [0.000, 0.000] - {counts[num_groups]} elements")
    for i in range(1, num_groups + 1):
        c = counts[i + num_groups]
        if c:
            if i == num_groups:
                print(f"      Bin {i:2d}: [{pos[i - 1]:.3f}, +inf) - {c} elements")
            else:
                print(f"      Bin {i:2d}: [{pos[i - 1]:.3f}, {pos[i]:.3f}) - {c} elements")
    print(f"      All {n} elements successfully assigned to bins")


# ---------------------------------------------------------------------------------------------------- vocabulary
def numerical_representation(data_points, tokens_out=None, integer_vocabulary=False):
    """data_utils.numerical_representation (data_utils.py:212-225): vocabulary = sorted(set(data)), token = index.
    List input: (list of int tokens, vocabulary list). Tensor input: (int32 device tokens — written into `tokens_out`, e.g.
    a row of the feeder's TokenStore — and the vocabulary as a Python list: it is what training_utils' metrics index).
    integer_vocabulary: the column holds group numbers (output of bin_numeric_data): vocabulary entries are ints."""
    as_list = not _is_column(data_points)
    if as_list:
        if not isinstance(data_points, list):
            raise TypeError("'data_points' must be a list.")
        if not data_points:
            return [], []
        if _first_non_number(data_points) >= 0:
            return _symbolic_representation(data_points, tokens_out)
        integer_vocabulary = integer_vocabulary or all(isinstance(v, int) for v in data_points)
    ops = _ops()
    x = _upload(data_points) if as_list else _check_column(data_points, "data_points")
    n = x.numel()
    dev = x.device
    st = _Status(ops, dev)
    counters = torch.empty(2, dtype=torch.int64, device=dev)
    capacity = 1 << 17
    limit = 1 << max(18, (2 * n - 1).bit_length())
    while True:
        table = torch.empty(capacity, dtype=torch.int64, device=dev)
        ops.prep_vocab_insert(x, table, counters, st.words)
        nonfinite, _, _, full = st.read()
        st.raise_nonfinite(nonfinite, x)
        distinct = int(counters[0])
        if full is None and 2 * distinct <= capacity:
            break
        if capacity >= limit:
            raise RuntimeError("vocabulary hash table did not converge")   # cannot happen: limit holds 2N slots
        capacity = min(limit, capacity * 16)
        ops.prep_status_reset(st.words)
    keys = torch.empty(distinct, dtype=torch.int64, device=dev)
    slots = torch.empty(distinct, dtype=torch.int32, device=dev)
    vocabulary = torch.empty(distinct, dtype=torch.float64, device=dev)
    token_of_slot = torch.empty(capacity, dtype=torch.int32, device=dev)
    ops.prep_vocab_build(table, distinct, keys, slots, vocabulary, token_of_slot, counters)
    tokens = torch.empty(n, dtype=torch.int32, device=dev) if tokens_out is None else tokens_out
    ops.prep_vocab_lookup(x, table, token_of_slot, tokens)
    vocab = vocabulary.tolist()
    if integer_vocabulary:
        vocab = [int(v) for v in vocab]
    return (tokens.tolist(), vocab) if as_list else (tokens, vocab)


def _symbolic_representation(data_points, tokens_out=None):
    """Non-numeric modalities (strings: a time-of-day or a label column; data_utils.py:212-225 takes any hashable, orderable
    values). The values exist only as host objects, so vocabulary = sorted(set(data)) is host work by nature and uses Python's
    own expressions (same order, same TypeError on values that do not compare); for a list of str the token ids come from
    numpy's sorted-unique inverse (code-point order, identical to Python's str ordering) instead of N dictionary look-ups.
    With `tokens_out` (a row of the device token store) the ids are uploaded there and that row is returned."""
    import numpy as np
    vocabulary = sorted(list(set(data_points)))
    if all(type(v) is str for v in data_points) and not any("\x00" in v for v in vocabulary):
        arr = np.array(data_points)               # fixed-width unicode; trailing NULs would be stripped, hence the check
        uniq, inverse = np.unique(arr, return_inverse=True)
        assert uniq.tolist() == vocabulary
        tokens = inverse.astype(np.int32)
    else:
        data_mapping = {element: index for index, element in enumerate(vocabulary)}
        tokens = np.fromiter((data_mapping[element] for element in data_points), dtype=np.int32, count=len(data_points))
    if tokens_out is None:
        return tokens.tolist(), vocabulary
    if tokens_out.numel() != tokens.size:
        raise ValueError(f"tokens_out holds {tokens_out.numel()} tokens, the modality has {tokens.size}")
    tokens_out.copy_(torch.from_numpy(tokens).to(tokens_out.device))
    return tokens_out, vocabulary


def create_train_val_datasets(numeric_rep_data, val_size, num_val_files, file_lengths):
    """data_utils.create_train_val_datasets (data_utils.py:228-290) for a device token row: returns two VIEWS of it
    (train, val) — same split arithmetic, same errors; no list is built."""
    if not torch.is_tensor(numeric_rep_data):
        raise TypeError("'numeric_rep_data' must be a device token tensor here (the reference's list form is unchanged).")
    if not isinstance(num_val_files, int) or num_val_files < 0:
        raise TypeError("'num_val_files' must be a non-negative integer.")
    if not isinstance(file_lengths, list) or not all(isinstance(v, int) and v > 0 for v in file_lengths):
        raise TypeError("'file_lengths' must be a list of positive integers.")
    n = int(numeric_rep_data.numel())
    if sum(file_lengths) != n:
        raise ValueError(f"Sum of file_lengths ({sum(file_lengths)}) does not match length of numeric_rep_data ({n}).")
    if num_val_files > 0:
        if num_val_files > len(file_lengths):
            raise ValueError(f"'num_val_files' ({num_val_files}) cannot exceed the number of loaded files "
                             f"({len(file_lengths)}).")
        train_num_elements = n - sum(file_lengths[-num_val_files:])
    else:
        if not isinstance(val_size, (int, float)) or not (0 < val_size < 1):
            raise ValueError("'val_size' must be a float between 0 and 1 (exclusive).")
        train_num_elements = int(n * (1 - val_size))
    return numeric_rep_data[:train_num_elements], numeric_rep_data[train_num_elements:]


# ---------------------------------------------------------------------------------------------------- the chain
def prepare_modality(column, modality_params, file_lengths=None, tokens_out=None, verbose=False, on_zero="keep"):
    """Raw column -> tokens, the steps main.py applies to one modality (main.py:100-210, 278) in the reference's order:
    percent changes (params[3]; per file, data_utils.py:148), ranging (params[4], params[5]), binning (params[6];
    outlier percentile params[10] default 0.1, exponent params[11] default 2.2, main.py:170-175), vocabulary.
    column: list of numbers or float64 CUDA tensor [N]. modality_params: the reference's legacy list (schema.py:213-248).
    on_zero: what a zero previous value does to the percent change - "keep" (default) is main.py's own loader,
    file_cache.load_file_data_cached (file_cache.py:366-372: 0.0 and a warning); "raise" is data_utils.load_file_data's
    ZeroDivisionError (data_utils.py:148-153).
    Returns (int32 device tokens [N], vocabulary list). Intermediate columns stay in HBM."""
    p = list(modality_params) + [None] * (12 - len(modality_params))
    is_percent, whole_digits, dec_places, num_bins = p[3], p[4], p[5], p[6]
    if not _is_column(column) and isinstance(column, list) and column and _first_non_number(column) >= 0 and \
            is_percent is not True and whole_digits is None and dec_places is None and num_bins is None:
        # a non-numeric modality without numeric processing steps (main.py applies none to such a column): vocabulary on the
        # host, token ids straight into the store row
        if tokens_out is None:
            tokens_out = torch.empty(len(column), dtype=torch.int32, device=_device())
        return _symbolic_representation(column, tokens_out)
    x = _upload(column) if not _is_column(column) else _check_column(column, "column")
    if is_percent is True:
        x = convert_to_percent_changes(x, dec_places if dec_places else 2, file_lengths=file_lengths, on_zero=on_zero)
    if whole_digits is not None or dec_places is not None:
        x = range_numeric_data(x, whole_digits, dec_places)
    binned = num_bins is not None
    if binned:
        outlier = p[10] if p[10] is not None else 0.1
        exponent = p[11] if p[11] is not None else 2.2
        x = bin_numeric_data(x, num_bins, outlier, exponent, verbose=verbose, as_float=True)
    return numerical_representation(x, tokens_out=tokens_out, integer_vocabulary=binned)


def prepare_dataset(columns, all_modality_params, file_lengths, validation_size=0.1, num_validation_files=0, verbose=False,
                    return_matrix=False):
    """All modalities of a run: what main.py builds between loading and training (main.py:100-210, 270-396), with the
    tokens kept in ONE int32 [M, N] device matrix — the layout of the feeder's TokenStore.
    columns: per modality a list of numbers or a float64 CUDA tensor [N] (raw values, files concatenated in order).
    Returns (all_train_sets, all_val_sets, all_vocabularies): per modality VIEWS of the matrix rows split as
    create_train_val_datasets splits them, and the vocabularies as Python lists — assign them to
    tradeaid_b200.training_utils' globals of the same names; get_batch then gathers straight from HBM.
    return_matrix=True appends the [M, N] matrix itself (what save_token_store takes)."""
    M = len(columns)
    if M == 0 or len(all_modality_params) != M:
        raise ValueError("columns and all_modality_params must list the same, non-zero number of modalities")
    lengths = {len(c) if not _is_column(c) else int(c.numel()) for c in columns}
    if len(lengths) != 1:
        raise ValueError(f"All modalities must have the same length for proper training, got {sorted(lengths)}.")
    n = lengths.pop()
    tokens = torch.empty(M, n, dtype=torch.int32, device=_device())
    vocabularies = []
    for i in range(M):
        _, vocab = prepare_modality(columns[i], all_modality_params[i], file_lengths, tokens_out=tokens[i], verbose=verbose)
        vocabularies.append(vocab)
    train, val = [], []
    for i in range(M):
        tr, va = create_train_val_datasets(tokens[i], validation_size, num_validation_files, list(file_lengths))
        train.append(tr)
        val.append(va)
    return (train, val, vocabularies, tokens) if return_matrix else (train, val, vocabularies)


# ---------------------------------------------------------------------------------------------------- token-store cache
# SURVEY.md 8 row f-4, second half: a prepared dataset is written once and mapped back on later runs, so that a 50 M-row
# run neither repeats the preparation nor builds Python lists. Layout (little endian):
#   bytes 0..7    magic  b"TATOKEN1"
#   bytes 8..15   uint64 length H of the JSON header
#   bytes 16..    JSON header: modalities M, rows N, file_lengths, all_modality_params, vocabularies (Python's float repr
#                 round-trips doubles exactly; group numbers stay ints), crc32 of the payload
#   then zero padding to a multiple of 4096 and the payload: int32 [M, N] row-major - the TokenStore matrix itself.
_TOKEN_MAGIC = b"TATOKEN1"
_TOKEN_ALIGN = 4096


def save_token_store(path, tokens, vocabularies, all_modality_params, file_lengths):
    """tokens: int32 [M, N] tensor (any device) or the list of its rows (e.g. prepare_dataset's train + val views are NOT
    the full rows - pass the matrix or full rows); vocabularies / all_modality_params / file_lengths as main.py holds them."""
    import json
    import zlib
    if not torch.is_tensor(tokens):
        tokens = torch.stack([t if torch.is_tensor(t) else torch.tensor(t, dtype=torch.int32) for t in tokens])
    if tokens.dim() != 2 or tokens.dtype != torch.int32:
        raise TypeError("tokens must be an int32 [M, N] matrix")
    if len(vocabularies) != tokens.shape[0] or len(all_modality_params) != tokens.shape[0]:
        raise ValueError("one vocabulary and one parameter list per modality")
    if sum(file_lengths) != tokens.shape[1]:
        raise ValueError(f"Sum of file_lengths ({sum(file_lengths)}) does not match the number of rows ({tokens.shape[1]}).")
    payload = tokens.contiguous().cpu().numpy().astype("<i4", copy=False)
    header = json.dumps({"modalities": int(tokens.shape[0]), "rows": int(tokens.shape[1]),
                         "file_lengths": [int(v) for v in file_lengths], "all_modality_params": [list(p) for p in all_modality_params],
                         "vocabularies": [list(v) for v in vocabularies], "crc32": zlib.crc32(payload.tobytes()) & 0xFFFFFFFF},
                        allow_nan=False).encode("utf-8")
    start = -(-(16 + len(header)) // _TOKEN_ALIGN) * _TOKEN_ALIGN
    tmp = f"{path}.tmp"
    with open(tmp, "wb") as f:
        f.write(_TOKEN_MAGIC)
        f.write(len(header).to_bytes(8, "little"))
        f.write(header)
        f.write(b"\0" * (start - 16 - len(header)))
        payload.tofile(f)
    import os
    os.replace(tmp, path)          # a reader never sees a half-written cache


def load_token_store(path, device=None, verify=True):
    """Returns (tokens int32 [M, N] on `device` (default: the current CUDA device), vocabularies, all_modality_params,
    file_lengths). The payload is memory-mapped and copied to the device in one transfer; verify=True checks its crc32."""
    import json
    import zlib
    with open(path, "rb") as f:
        if f.read(8) != _TOKEN_MAGIC:
            raise ValueError(f"{path} is not a tradeaid_b200 token-store cache")
        hlen = int.from_bytes(f.read(8), "little")
        header = json.loads(f.read(hlen).decode("utf-8"))
    M, N = header["modalities"], header["rows"]
    start = -(-(16 + hlen) // _TOKEN_ALIGN) * _TOKEN_ALIGN
    import os
    if os.path.getsize(path) != start + 4 * M * N:
        raise ValueError(f"{path}: truncated or oversized payload (expected {start + 4 * M * N} bytes)")
    payload = np.memmap(path, dtype="<i4", mode="r", offset=start, shape=(M, N))
    if verify and (zlib.crc32(payload.tobytes()) & 0xFFFFFFFF) != header["crc32"]:
        raise ValueError(f"{path}: payload checksum mismatch")
    dev = torch.device(device) if device is not None else _device()
    tokens = torch.from_numpy(np.array(payload, dtype=np.int32)).to(dev)     # (a private copy: the mapping is read-only)
    return tokens, header["vocabularies"], header["all_modality_params"], header["file_lengths"]
