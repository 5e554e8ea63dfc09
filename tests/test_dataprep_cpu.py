"""Row f-4 (one-shot data preparation) without a GPU:
  * the oracle restatement (oracle/dataprep.py) against the fixtures made by the reference's data_utils
    (tests/golden/dataprep.pt, generator tests/golden/make_golden_dataprep.py) — bit for bit;
  * the element arithmetic the CUDA kernels run (csrc/dataprep_math.cuh), compiled for the host with g++
    (tests/host_harness/dataprep_host.cpp), against CPython and the same fixtures — bit for bit;
  * the host scalars of tradeaid_b200.data_utils (libm decade thresholds, 10**k factors, numpy's percentile
    interpolation, group boundaries) and the argument validation / error messages of the drop-in functions."""
import ctypes
import math
import os
import random
import struct
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from oracle import dataprep as oracle

P = ctypes.c_void_p


def bits(values):
    return np.ascontiguousarray(values, dtype=np.float64).view(np.uint64)


def same_bits(a, b):
    """Equal as doubles bit for bit (the sign of zero included)."""
    return np.array_equal(bits(a), bits(b))


@pytest.fixture(scope="module")
def golden():
    return load_golden("dataprep.pt")


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    """The device arithmetic built for the host."""
    out = str(tmp_path_factory.mktemp("dph") / "libdataprep_host.so")
    src = os.path.join(ROOT, "tests", "host_harness", "dataprep_host.cpp")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, src], check=True)
    return ctypes.CDLL(out)


def col(g, case):
    return g["columns"][case["x"]].tolist()


# ------------------------------------------------------------------------------------------------ oracle vs reference
def test_oracle_percent_changes(golden):
    for c in golden["percent"]:
        x = col(golden, c)
        if "file_lengths" in c:
            y = oracle.percent_changes_by_file(x, c["file_lengths"], c["decimals"])
        else:
            y = oracle.percent_changes(x, c["decimals"], c.get("on_zero", "raise"))
        assert same_bits(y, c["y"].numpy()), c["name"]
    e = golden["errors"]["zero_prev"]
    with pytest.raises(ZeroDivisionError) as info:
        oracle.percent_changes(col(golden, e), 2)
    assert str(info.value) == e["message"]


def test_oracle_range(golden):
    for c in golden["range"]:
        y = oracle.range_values(col(golden, c), c["whole_digits"], c["decimals"])
        assert same_bits([float(v) for v in y], c["y"].numpy()), (c["name"], c["whole_digits"], c["decimals"])


def test_oracle_bins(golden):
    for c in golden["bin"]:
        y = oracle.bin_groups(col(golden, c), c["num_groups"], c["outlier"], c["exponent"])
        assert y == c["y"].tolist(), (c["name"], c["num_groups"], c["outlier"], c["exponent"])
    e = golden["errors"]["all_filtered"]
    with pytest.raises(ValueError) as info:
        oracle.bin_groups(col(golden, e), e["num_groups"], e["outlier"], e["exponent"])
    assert str(info.value) == e["message"]


def test_oracle_vocabulary_and_chain(golden):
    for c in golden["vocab"]:
        tok, voc = oracle.vocabulary_tokens(col(golden, c))
        assert tok == c["tokens"].tolist() and voc == c["vocabulary"].tolist(), c["name"]
    for c in golden["chain"]:
        tok, voc = oracle.prepare_modality(col(golden, c), c["params"], c["file_lengths"])
        assert tok == c["tokens"].tolist(), c["name"]
        assert [float(v) for v in voc] == c["vocabulary"].tolist(), c["name"]
        assert all(isinstance(v, int) for v in voc) == c["integer_vocabulary"]


# ------------------------------------------------------------------------------------------------ device arithmetic on the host
def h_round(host, xs, nd):
    x = np.ascontiguousarray(xs, dtype=np.float64)
    o = np.empty_like(x)
    host.h_py_round(P(x.ctypes.data), ctypes.c_longlong(x.size), ctypes.c_int(nd), P(o.ctypes.data))
    return o


def test_device_round_is_cpython_round(host):
    rng = random.Random(7)
    for nd in range(0, 23):
        xs = [0.0, -0.0, 0.5, 1.5, 2.5, -0.5, 0.125, 0.375, 2.675, 1.005, 1e15 + 0.5, 4503599627370495.5, 5e-324, 1e22, 0.1,
              -2.5e-3, 2 ** 52 + 0.0, 2 ** 53 + 2.0, 1e300, -1e-300]
        for _ in range(3000):
            c = rng.random()
            if c < 0.3:
                xs.append(rng.uniform(-1000, 1000))
            elif c < 0.5:
                xs.append(rng.randint(-10 ** 7, 10 ** 7) / 2 ** rng.randint(0, 30))          # exact ties
            elif c < 0.6:
                xs.append((rng.randint(-10 ** 6, 10 ** 6) + 0.5) / 10 ** rng.randint(0, 8))
            elif c < 0.75:
                xs.append(rng.uniform(-1, 1) * 10 ** rng.randint(-30, 15))
            elif c < 0.85:
                v = struct.unpack("d", struct.pack("Q", rng.getrandbits(64)))[0]
                if v == v and abs(v) != math.inf:
                    xs.append(v)
            else:
                xs.append(round(rng.uniform(-100, 100), rng.randint(0, 6)))
        want = []
        keep = []
        for v in xs:
            try:
                want.append(round(v, nd))
                keep.append(v)
            except OverflowError:      # CPython refuses results beyond DBL_MAX; the data path never gets there
                pass
        assert same_bits(h_round(host, keep, nd), want), nd


def test_division_free_quotient_is_ieee_division(host):
    """k / 10^nd by Markstein's correction with a host-made reciprocal (csrc/dataprep_math.cuh div_pow10) equals the IEEE
    quotient for every sampled integer k <= 2^53 (all bit lengths, the smallest and the largest 4096 values)."""
    host.h_div_pow10_mismatches.restype = ctypes.c_longlong
    for nd in range(0, 23):
        assert host.h_div_pow10_mismatches(ctypes.c_int(nd), ctypes.c_longlong(2_000_000), ctypes.c_ulonglong(1234 + nd)) == 0, nd


def thresholds_and_factors(nwd):
    from tradeaid_b200 import data_utils as du
    return (np.array(du._log10_thresholds(), dtype=np.float64),
            np.array(du._decade_factors(nwd if nwd is not None else 1), dtype=np.float64))


def test_decade_thresholds_follow_libm(host):
    from tradeaid_b200 import data_utils as du
    thr = du._log10_thresholds()
    assert len(thr) == 633 and all(a < b for a, b in zip(thr, thr[1:]))
    rng = random.Random(3)
    xs = []
    for k, t in enumerate(thr):
        p = k - 324
        assert math.floor(math.log10(t)) >= p
        if t > 5e-324:
            assert math.floor(math.log10(math.nextafter(t, -math.inf))) < p
        v = t
        for _ in range(24):             # a neighbourhood of every decade boundary, both sides
            xs.append(v)
            v = math.nextafter(v, math.inf)
        v = t
        for _ in range(24):
            v = math.nextafter(v, -math.inf)
            if v > 0:
                xs.append(v)
    xs += [rng.uniform(0, 1) * 10 ** rng.randint(-300, 300) for _ in range(20000)]
    xs = [v for v in xs if v > 0]
    x = np.array(xs, dtype=np.float64)
    t = np.array(thr, dtype=np.float64)
    out = np.empty(len(xs), dtype=np.int32)
    host.h_floor_log10(P(x.ctypes.data), ctypes.c_longlong(x.size), P(t.ctypes.data), P(out.ctypes.data))
    assert out.tolist() == [int(math.floor(math.log10(v))) for v in xs]


def test_device_range_against_reference(host, golden):
    from tradeaid_b200 import data_utils as du
    for c in golden["range"]:
        nwd, dp = c["whole_digits"], c["decimals"]
        x = np.ascontiguousarray(golden["columns"][c["x"]].numpy())
        thr, fac = thresholds_and_factors(nwd)
        lower = upper = clamp = 0.0
        if nwd is not None:
            lower, upper = float(10 ** (nwd - 1)), float(10 ** nwd)
            clamp = float(10 ** nwd - (10 ** (-dp))) if dp > 0 else float(10 ** nwd - 1)
        out = np.empty_like(x)
        host.h_range(P(x.ctypes.data), ctypes.c_longlong(x.size), ctypes.c_int(nwd is not None), ctypes.c_int(dp),
                     ctypes.c_double(lower), ctypes.c_double(upper), ctypes.c_double(clamp), P(thr.ctypes.data),
                     P(fac.ctypes.data), P(out.ctypes.data))
        want = c["y"].numpy()
        assert same_bits(out, want), (c["name"], nwd, dp)
    assert du._decade_factors(1)[324] == 1.0 and du._decade_factors(3)[324 + 5] == 10.0 ** -3


def test_device_percent_and_bins_against_reference(host, golden):
    for c in golden["percent"]:
        if "file_lengths" in c:
            continue
        x = np.ascontiguousarray(golden["columns"][c["x"]].numpy())
        if c.get("on_zero") != "keep" and (x[:-1] == 0).any():
            continue
        out = np.empty_like(x)
        host.h_percent(P(x.ctypes.data), ctypes.c_longlong(x.size), ctypes.c_int(2 if c["decimals"] is None else c["decimals"]),
                       P(out.ctypes.data))
        assert same_bits(out, c["y"].numpy()), (c["name"], c["decimals"])
    # unchanged values (answered without a division): zero of the previous value's sign, like Python's 0.0 / prev
    x = np.array([3.0, 3.0, -2.0, -2.0, 7.5, 7.5, -1e-300, -1e-300, 4.0], dtype=np.float64)
    out = np.empty_like(x)
    host.h_percent(P(x.ctypes.data), ctypes.c_longlong(x.size), ctypes.c_int(2), P(out.ctypes.data))
    assert same_bits(out, oracle.percent_changes(x.tolist(), 2))
    for c in golden["bin"]:
        x = np.ascontiguousarray(golden["columns"][c["x"]].numpy())
        lo, hi = np.percentile(x, c["outlier"]), np.percentile(x, 100 - c["outlier"])
        kept = x[(x >= lo) & (x <= hi)]
        pos, neg = oracle.group_boundaries(c["num_groups"], c["exponent"], max(abs(kept.min()), abs(kept.max())))
        pos, neg = np.array(pos), np.array(neg)
        assert np.all(np.diff(pos) >= 0) and np.all(np.diff(neg) >= 0)
        for is_sorted in (0, 1):                                  # the reference's linear walk and the bisection
            out = np.empty(x.size, dtype=np.int32)
            host.h_bin(P(x.ctypes.data), ctypes.c_longlong(x.size), P(pos.ctypes.data), P(neg.ctypes.data),
                       ctypes.c_int(c["num_groups"]), ctypes.c_int(is_sorted), P(out.ctypes.data))
            assert out.tolist() == c["y"].tolist(), (c["name"], c["num_groups"], is_sorted)
    # boundaries with repeats (empty groups): both searches agree on every side of every boundary
    pos = np.array([0.0, 1.0, 1.0, 2.0, 2.0, 2.0, 5.0]); neg = -pos[::-1].copy(); neg[-1] = 0.0
    xs = np.array(sorted({s * (b + d) for b in pos for d in (-0.5, -1e-9, 0.0, 1e-9, 0.5) for s in (1, -1)} | {7.0, -7.0}))
    outs = []
    for is_sorted in (0, 1):
        out = np.empty(xs.size, dtype=np.int32)
        host.h_bin(P(xs.ctypes.data), ctypes.c_longlong(xs.size), P(pos.ctypes.data), P(neg.ctypes.data), ctypes.c_int(6),
                   ctypes.c_int(is_sorted), P(out.ctypes.data))
        outs.append(out.tolist())
    assert outs[0] == outs[1]


def test_order_key_is_monotone_and_invertible(host):
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.standard_normal(5000) * 10.0 ** rng.integers(-20, 20, 5000), [0.0, -0.0, 5e-324, -5e-324, 1e308, -1e308]])
    keys = np.empty(x.size, dtype=np.uint64)
    back = np.empty_like(x)
    host.h_order_key(P(x.ctypes.data), ctypes.c_longlong(x.size), P(keys.ctypes.data), P(back.ctypes.data))
    assert same_bits(back, x)
    order = np.argsort(keys, kind="stable")
    assert np.all(np.diff(x[order]) >= 0)


# ------------------------------------------------------------------------------------------------ host scalars
def test_percentile_plan_is_numpys():
    from tradeaid_b200 import data_utils as du
    rng = np.random.default_rng(11)
    for n in [1, 2, 3, 4, 5, 10, 11, 97, 1000, 1501]:
        a = np.sort(rng.standard_normal(n) * 100)
        shuffled = rng.permutation(a)
        for q in [0, 0.1, 1, 5, 12.5, 25, 45, 50, 55, 75, 87.5, 95, 99, 99.9, 100, 100 - 0.1, 100 - 12.5]:
            prev, nxt, gamma = du._percentile_plan(n, q)
            got = du._lerp(float(a[prev]), float(a[nxt]), gamma)
            want = float(np.percentile(list(shuffled), q))
            assert struct.pack("d", got) == struct.pack("d", want), (n, q)


def test_group_boundaries_match_oracle():
    from tradeaid_b200 import data_utils as du
    for G, ex, m in [(5, 2.0, 3.7), (10, 2.2, 0.123), (64, 1, 1e6), (3, 3.0, 0.0)]:
        assert du._group_boundaries(G, ex, m) == oracle.group_boundaries(G, ex, m)


def test_validation_matches_reference_messages():
    """Argument checks run before anything touches the device: same exception types and texts as data_utils.py."""
    from tradeaid_b200 import data_utils as du
    with pytest.raises(ValueError, match="'data' must be a non-empty list."):
        du.convert_to_percent_changes([], 2)
    with pytest.raises(ValueError, match="Element at index 1 is str: 'a'"):
        du.convert_to_percent_changes([1.0, "a"], 2)
    with pytest.raises(ValueError, match="'decimal_places' must be a non-negative integer or null."):
        du.convert_to_percent_changes([1.0, 2.0], -1)
    with pytest.raises(TypeError, match="'numeric_data' must be a list."):
        du.range_numeric_data((1.0, 2.0), 1, 2)
    with pytest.raises(TypeError, match="'numeric_data' must be a non-empty list."):
        du.range_numeric_data([], 1, 2)
    with pytest.raises(IndexError, match="Element at index 0 in 'numeric_data' is not a number."):
        du.range_numeric_data(["x"], 1, 2)
    with pytest.raises(TypeError, match="'num_whole_digits' must be an integer or None."):
        du.range_numeric_data([1.0], 1.5, 2)
    with pytest.raises(ValueError, match="'decimal_places' must be greater than or equal to 0."):
        du.range_numeric_data([1.0], 1, -2)
    with pytest.raises(NotImplementedError):
        du.range_numeric_data([1.0], 1, 40)
    with pytest.raises(ValueError, match="'num_groups' must be a positive integer."):
        du.bin_numeric_data([1.0], 0)
    with pytest.raises(ValueError, match="'outlier_percentile' must be a number between 0 and 50."):
        du.bin_numeric_data([1.0], 3, 60)
    with pytest.raises(ValueError, match="'exponent' must be a number >= 1."):
        du.bin_numeric_data([1.0], 3, 5, 0.5)
    with pytest.raises(TypeError, match="'file_lengths' must be a list of positive integers."):
        du.create_train_val_datasets(torch.zeros(4, dtype=torch.int32), 0.1, 0, [4, 0])
    with pytest.raises(ValueError, match="'val_size' must be a float between 0 and 1"):
        du.create_train_val_datasets(torch.zeros(4, dtype=torch.int32), 1.5, 0, [4])
    tr, va = du.create_train_val_datasets(torch.arange(10, dtype=torch.int32), 0.1, 0, [4, 6])
    assert tr.tolist() == list(range(9)) and va.tolist() == [9]          # int(10 * (1 - 0.1)) = 9 (data_utils.py:276)
    tr, va = du.create_train_val_datasets(torch.arange(10, dtype=torch.int32), 0.1, 1, [4, 6])
    assert tr.numel() == 4 and va.numel() == 6


def test_no_cpu_path():
    """The drop-in functions never compute on the host: without a CUDA device they raise."""
    from tradeaid_b200 import data_utils as du
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    for call in (lambda: du.convert_to_percent_changes([1.0, 2.0], 2), lambda: du.range_numeric_data([1.0, 2.0], 1, 2),
                 lambda: du.bin_numeric_data([1.0, 2.0], 3), lambda: du.numerical_representation([1.0, 2.0])):
        with pytest.raises(RuntimeError, match="no CPU path"):
            call()


# ------------------------------------------------------------------------------------------------ host logic over the CPU emulation
@pytest.fixture()
def du_emu():
    """tradeaid_b200.data_utils with the K9 kernel namespace replaced by tests/emu_prep.py (the kernels' own element
    arithmetic built for the host + numpy data movement): everything data_utils does on the host runs for real."""
    import emu_prep
    from tradeaid_b200 import data_utils as du
    du._ops_override = emu_prep
    try:
        yield du
    finally:
        du._ops_override = None


def test_host_logic_percent_range_bins_golden(du_emu, golden, capsys):
    du = du_emu
    for c in golden["percent"]:
        x = golden["columns"][c["x"]]
        y = du.convert_to_percent_changes(x.clone(), c["decimals"], file_lengths=c.get("file_lengths"),
                                          on_zero=c.get("on_zero", "raise"))
        assert same_bits(y.numpy(), c["y"].numpy()), (c["name"], c["decimals"])
    e = golden["errors"]["zero_prev"]
    with pytest.raises(ZeroDivisionError) as info:
        du.convert_to_percent_changes(golden["columns"][e["x"]].tolist(), 2)
    assert str(info.value) == e["message"]
    for c in golden["range"]:
        y = du.range_numeric_data(golden["columns"][c["x"]].tolist(), c["whole_digits"], c["decimals"])
        assert isinstance(y, list) and same_bits(y, c["y"].numpy()), (c["name"], c["whole_digits"], c["decimals"])
    e = golden["errors"]["range_overflow"]
    with pytest.raises(OverflowError):
        du.range_numeric_data(golden["columns"][e["x"]].tolist(), e["whole_digits"], e["decimals"])
    with pytest.raises(ValueError, match="non-finite value nan at index 1"):
        du.range_numeric_data([1.0, float("nan"), 2.0], 1, 2)
    for c in golden["bin"]:
        y = du.bin_numeric_data(golden["columns"][c["x"]].clone(), c["num_groups"], c["outlier"], c["exponent"], verbose=False)
        assert y.dtype == torch.int64 and y.tolist() == c["y"].tolist(), (c["name"], c["num_groups"], c["outlier"])
    e = golden["errors"]["all_filtered"]
    with pytest.raises(ValueError) as info:
        du.bin_numeric_data(golden["columns"][e["x"]].tolist(), e["num_groups"], e["outlier"], e["exponent"])
    assert str(info.value) == e["message"]
    capsys.readouterr()
    c = golden["bin"][0]
    du.bin_numeric_data(golden["columns"][c["x"]].tolist(), c["num_groups"], c["outlier"], c["exponent"])
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.strip().startswith("Bin")]
    assert len(lines) == len(set(c["y"].tolist()))


def test_host_logic_vocabulary_chain_and_split(du_emu, golden):
    import emu_prep
    du = du_emu
    for c in golden["vocab"]:
        tok, voc = du.numerical_representation(golden["columns"][c["x"]].clone())
        assert tok.dtype == torch.int32 and tok.tolist() == c["tokens"].tolist() and voc == c["vocabulary"].tolist(), c["name"]
    assert du.numerical_representation([3, 1, 2, 3, 1]) == ([2, 0, 1, 2, 0], [1, 2, 3])
    for c in golden["chain"]:
        x = golden["columns"][c["x"]]
        row = torch.full((x.numel(),), -7, dtype=torch.int32)
        tok, voc = du.prepare_modality(x.tolist(), c["params"], c["file_lengths"], tokens_out=row)
        assert tok is row and row.tolist() == c["tokens"].tolist(), c["name"]
        assert [float(v) for v in voc] == c["vocabulary"].tolist()
        assert all(isinstance(v, int) for v in voc) == c["integer_vocabulary"]
    # a zero inside a percent modality: main.py's loader keeps 0.0 (default), data_utils' loader raises
    zc = [5.0, 0.0, 3.0, 3.3, 0.0, 0.0, 1.0]
    zp = ["f", 1, None, True, None, 2, None, None, False, "z", None, None]
    tok, voc = du.prepare_modality(zc, zp)
    want_tok, want_voc = oracle.prepare_modality(zc, zp)
    assert tok.tolist() == want_tok and voc == want_voc and 0.0 in voc
    with pytest.raises(ZeroDivisionError, match="previous value is zero at index 1"):
        du.prepare_modality(zc, zp, on_zero="raise")
    # ... and the reference converts file by file, so its message carries the index inside the file
    with pytest.raises(ZeroDivisionError) as info:
        du.convert_to_percent_changes([1.0, 2.0, 3.0, 4.0, 0.0, 6.0, 0.0, 1.0], 2, file_lengths=[3, 5])
    with pytest.raises(ZeroDivisionError) as want:
        oracle.percent_changes_by_file([1.0, 2.0, 3.0, 4.0, 0.0, 6.0, 0.0, 1.0], [3, 5], 2)
    assert str(info.value) == str(want.value) == "Cannot calculate percentage change: previous value is zero at index 1."
    # more distinct values than half the first table: one retry with a larger table, same result
    before = emu_prep.stats["vocab_inserts"]
    n = 70_000
    x = torch.from_numpy(np.random.default_rng(0).permutation(n).astype(np.float64) * 0.25 - 100.0)
    tok, voc = du.numerical_representation(x)
    assert emu_prep.stats["vocab_inserts"] - before == 2
    assert len(voc) == n and voc[0] == -100.0 and torch.equal(tok.long(), ((x + 100.0) * 4).long())
    # all modalities into one [M, N] matrix + the reference's split arithmetic
    a, b = golden["chain"][0], golden["chain"][3]
    cols = [golden["columns"][a["x"]].tolist(), golden["columns"][b["x"]].clone()]
    n = len(cols[0])
    train, val, vocabs = du.prepare_dataset(cols, [a["params"], b["params"]], [n], validation_size=0.2)
    cut = int(n * (1 - 0.2))
    for i, c in enumerate((a, b)):
        assert train[i].tolist() == c["tokens"].tolist()[:cut] and val[i].tolist() == c["tokens"].tolist()[cut:]
        assert [float(v) for v in vocabs[i]] == c["vocabulary"].tolist()
    assert train[0].data_ptr() + 4 * n == train[1].data_ptr()          # rows of one matrix
    matrix = du.prepare_dataset(cols, [a["params"], b["params"]], [n], validation_size=0.2, return_matrix=True)[3]
    assert matrix.shape == (2, n) and matrix[0].tolist() == a["tokens"].tolist() and matrix[1].tolist() == b["tokens"].tolist()
    with pytest.raises(ValueError, match="same length"):
        du.prepare_dataset([[1.0, 2.0], [1.0]], [a["params"], b["params"]], [2])


def test_reference_demo_columns(du_emu):
    """BASELINE.json configs[0]: the reference's own demo schemas on its demo_stock.csv (tests/golden/demo_prep.pt, made by
    the reference's loader + processing functions in main.py's order): vocabulary sizes [57, 3], tokens and vocabularies
    reproduced by the oracle and by data_utils' host logic over the kernels' arithmetic."""
    g = load_golden("demo_prep.pt")
    assert [v.numel() for v in g["vocabulary"]] == [57, 3]
    for raw, params, tokens, vocabulary in zip(g["raw"], g["params"], g["tokens"], g["vocabulary"]):
        tok, voc = oracle.prepare_modality(raw.tolist(), params, g["file_lengths"])
        assert tok == tokens.tolist() and [float(v) for v in voc] == vocabulary.tolist()
        tok, voc = du_emu.prepare_modality(raw.tolist(), params, g["file_lengths"])
        assert tok.tolist() == tokens.tolist() and [float(v) for v in voc] == vocabulary.tolist()


def test_token_store_cache_round_trip(tmp_path):
    """save_token_store / load_token_store (row f-4's binary cache): tokens, vocabularies (doubles bit for bit, group
    numbers as ints, strings), parameters and file lengths come back unchanged; corruption and truncation are detected."""
    from tradeaid_b200 import data_utils as du
    g = torch.Generator().manual_seed(2)
    tokens = torch.randint(0, 1000, (3, 5000), generator=g, dtype=torch.int32)
    vocabs = [[0.1 * k - 3.0 for k in range(1000)] + [-0.0, 1e-300, 123456789.125], list(range(-8, 9)), [f"sym{k}" for k in range(5)]]
    params = [["a.csv", 1, True, True, None, 2, 8, None, False, "pct", 0.1, 2.2], ["a.csv", 2, True, False, 2, 1, None, None, True, "rng", None, None],
              ["a.csv", 3, True, False, None, None, None, None, False, "sym", None, None]]
    path = str(tmp_path / "store.tatok")
    du.save_token_store(path, tokens, vocabs, params, [2000, 3000])
    tok, voc, prm, lens = du.load_token_store(path, device="cpu")
    assert tok.dtype == torch.int32 and torch.equal(tok, tokens) and prm == params and lens == [2000, 3000]
    assert voc[1] == vocabs[1] and all(isinstance(v, int) for v in voc[1]) and voc[2] == vocabs[2]
    assert same_bits(voc[0], vocabs[0])
    du.save_token_store(path, [tokens[i] for i in range(3)], vocabs, params, [5000])        # rows instead of the matrix
    assert torch.equal(du.load_token_store(path, device="cpu")[0], tokens)
    with pytest.raises(ValueError, match="Sum of file_lengths"):
        du.save_token_store(path, tokens, vocabs, params, [1, 2])
    raw = bytearray(open(path, "rb").read())
    raw[-5] ^= 0x40
    open(path, "wb").write(raw)
    with pytest.raises(ValueError, match="checksum"):
        du.load_token_store(path, device="cpu")
    assert du.load_token_store(path, device="cpu", verify=False)[0].shape == (3, 5000)
    open(path, "wb").write(raw[:-8])
    with pytest.raises(ValueError, match="truncated"):
        du.load_token_store(path, device="cpu")
    open(path, "wb").write(b"not a cache file at all")
    with pytest.raises(ValueError, match="not a tradeaid_b200"):
        du.load_token_store(path, device="cpu")


def test_string_modalities_are_tokenised_like_the_reference(du_emu):
    """numerical_representation on non-numeric data (data_utils.py:212-225: vocabulary = sorted(set(data)), token = index) —
    the case that used to raise NotImplementedError: same vocabulary order (Python's / code-point order), same tokens, same
    TypeError for values that do not compare; through prepare_modality the ids land in the token-store row."""
    import random
    du = du_emu
    rng = random.Random(5)
    words = ["09:30", "09:35", "10:00", "close", "Open", "open", "é", "zz", "10:0", "", " "]
    data = [rng.choice(words) for _ in range(500)]
    tokens, vocab = du.numerical_representation(data)
    want_vocab = sorted(list(set(data)))
    mapping = {e: i for i, e in enumerate(want_vocab)}
    assert vocab == want_vocab and tokens == [mapping[e] for e in data]
    mixed = [("a", 1), ("a", 0), ("b", 2)] * 3                      # tuples: hashable and orderable, not str
    t2, v2 = du.numerical_representation(mixed)
    assert v2 == sorted(set(mixed)) and t2 == [v2.index(e) for e in mixed]
    with pytest.raises(TypeError):                                  # sorted() of str and None, as in the reference
        du.numerical_representation(["a", None, "b"])
    row = torch.full((len(data),), -1, dtype=torch.int32)
    params = ["t.csv", 9, True, False, None, None, None, None, True, "Time of day", None, None]
    tok, voc = du.prepare_modality(data, params, tokens_out=row)
    assert tok is row and voc == want_vocab and row.tolist() == tokens
    tr, va, vo = du.prepare_dataset([data, [float(len(w)) for w in data]], [params, ["t.csv", 1, None, False, None, None, None, None,
                                                                                       False, "len", None, None]], [500], 0.2)
    assert vo[0] == want_vocab and torch.cat([tr[0], va[0]]).tolist() == tokens


def test_repr_decimals_is_the_length_of_str_fraction(host):
    """The device's count of decimals for decimal_places=None (smallest d at which round() reproduces x, at least 1) against
    Python's own len(str(x).split('.')[-1]) on prices, random doubles of every positional magnitude, and short decimals."""
    import random
    rng = random.Random(11)
    vals = [0.0, -0.0, 1.0, 100.0, 0.1, 0.5, 12.5, 0.125, 1e-4, 9999999999999998.0, 123456789012345.6, 0.0001234, 5e-324 * 0 + 0.3,
            1 / 3, 2 / 3, 1e15 + 0.5, 4.35, 2.675, 1.005, 0.30000000000000004]
    for _ in range(60000):
        k = rng.random()
        if k < 0.3:
            vals.append(round(rng.uniform(0.01, 5000.0), rng.randint(0, 6)))          # prices / quoted decimals
        elif k < 0.6:
            vals.append(rng.uniform(-1, 1) * 10.0 ** rng.randint(-4, 15))             # full-precision doubles
        elif k < 0.8:
            vals.append(rng.randint(-10 ** 6, 10 ** 6) / 10.0 ** rng.randint(0, 8))
        else:
            vals.append(float(rng.randint(-10 ** 15, 10 ** 15)))                       # integral floats: "x.0"
    vals = [v for v in vals if v == 0.0 or 1e-4 <= abs(v) < 1e16]
    x = np.asarray(vals, dtype=np.float64)
    out = np.zeros(x.size, dtype=np.int32)
    host.h_repr_decimals(ctypes.c_void_p(x.ctypes.data), ctypes.c_longlong(x.size), ctypes.c_void_p(out.ctypes.data))
    want = [len(str(float(v)).split('.')[-1]) for v in x.tolist()]
    assert all('e' not in str(float(v)) for v in x.tolist())
    bad = [(v, int(o), w) for v, o, w in zip(x.tolist(), out.tolist(), want) if o != w]
    assert not bad, bad[:5]
    # exponent notation is left to the host
    y = np.asarray([1e16, 1.5e-5, 1e-5, 2.5e22, -3e-7], dtype=np.float64)
    o2 = np.zeros(y.size, dtype=np.int32)
    host.h_repr_decimals(ctypes.c_void_p(y.ctypes.data), ctypes.c_longlong(y.size), ctypes.c_void_p(o2.ctypes.data))
    assert o2.tolist() == [-1] * 5


def test_range_with_decimal_places_none_matches_the_reference_algorithm(du_emu):
    """range_numeric_data(decimal_places=None) through the host wrapper over the kernels' arithmetic == the oracle's restatement
    of data_utils.py:421-457 (and the reference itself where its checkout exists), incl. ints, exponent-notation values."""
    import random
    from oracle import dataprep as O
    du = du_emu
    rng = random.Random(3)
    data = [12.5, 0.125, 100.0, 7, -3, 0.0, 1.5e-5, 2e16, -0.00042, 99.999, 1234.5678, 0.30000000000000004, 5, 1e-4]
    data += [round(rng.uniform(-3000, 3000), rng.randint(0, 5)) for _ in range(3000)]
    data += [rng.uniform(-1, 1) * 10.0 ** rng.randint(-6, 17) for _ in range(500)]
    for nwd in (None, 1, 2, 3):
        got = du.range_numeric_data(list(data), nwd, None)
        want = O.range_values(data, nwd, None)
        assert len(got) == len(want)
        bad = [(v, g, w) for v, g, w in zip(data, got, want) if g != w or math.copysign(1, g) != math.copysign(1, w)]
        assert not bad, (nwd, bad[:5])
    col = torch.tensor([v for v in data if isinstance(v, float)], dtype=torch.float64)
    got = du.range_numeric_data(col, 2, None)
    assert got.tolist() == O.range_values(col.tolist(), 2, None)
    if os.path.isdir("/root/reference"):
        import importlib.util
        spec = importlib.util.spec_from_file_location("_ref_du", "/root/reference/data_utils.py")
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        assert ref.range_numeric_data(list(data), 2, None) == O.range_values(data, 2, None)
