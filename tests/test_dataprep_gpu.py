"""Row f-4 on the device: the K9 kernels (csrc/dataprep.cu) through the C ABI and the drop-in functions of
tradeaid_b200.data_utils, bit-exact against the reference's fixtures (tests/golden/dataprep.pt), against the oracle
(oracle/dataprep.py) on fresh seeded columns, and at full size through properties (vocabulary round trip, order
statistics against a sort, percent changes against float64 torch arithmetic + the CPU harness-checked round)."""
import math
import random

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def bits(values):
    return np.ascontiguousarray(values, dtype=np.float64).view(np.uint64)


def same_bits(a, b):
    return np.array_equal(bits(a), bits(b))


@pytest.fixture(scope="module")
def golden():
    return load_golden("dataprep.pt")


@pytest.fixture(scope="module")
def du():
    from tradeaid_b200 import data_utils
    return data_utils


def dev(t):
    return t.to("cuda")


# ------------------------------------------------------------------------------------------------ against the reference
def test_percent_changes_golden(du, golden):
    for c in golden["percent"]:
        x = golden["columns"][c["x"]]
        kw = dict(file_lengths=c.get("file_lengths"), on_zero=c.get("on_zero", "raise"))
        y = du.convert_to_percent_changes(dev(x), c["decimals"], **kw)
        assert y.is_cuda and same_bits(y.cpu().numpy(), c["y"].numpy()), (c["name"], c["decimals"])
    c = golden["percent"][0]
    y = du.convert_to_percent_changes(golden["columns"][c["x"]].tolist(), c["decimals"])     # list in, list out
    assert isinstance(y, list) and same_bits(y, c["y"].numpy())
    e = golden["errors"]["zero_prev"]
    with pytest.raises(ZeroDivisionError) as info:
        du.convert_to_percent_changes(golden["columns"][e["x"]].tolist(), 2)
    assert str(info.value) == e["message"]


def test_range_golden(du, golden):
    for c in golden["range"]:
        x = dev(golden["columns"][c["x"]])
        y = du.range_numeric_data(x, c["whole_digits"], c["decimals"])
        assert same_bits(y.cpu().numpy(), c["y"].numpy()), (c["name"], c["whole_digits"], c["decimals"])
    c = golden["range"][0]
    x = dev(golden["columns"][c["x"]]).clone()
    assert du.range_numeric_data(x, c["whole_digits"], c["decimals"], out=x) is x          # in place
    assert same_bits(x.cpu().numpy(), c["y"].numpy())
    y = du.range_numeric_data(golden["columns"][c["x"]].tolist(), c["whole_digits"], c["decimals"])
    assert isinstance(y, list) and same_bits(y, c["y"].numpy())
    e = golden["errors"]["range_overflow"]
    with pytest.raises(OverflowError):
        du.range_numeric_data(golden["columns"][e["x"]].tolist(), e["whole_digits"], e["decimals"])
    with pytest.raises(ValueError, match="non-finite"):
        du.range_numeric_data([1.0, float("nan"), 2.0], 1, 2)


def test_bins_golden(du, golden, capsys):
    for c in golden["bin"]:
        x = dev(golden["columns"][c["x"]])
        y = du.bin_numeric_data(x, c["num_groups"], c["outlier"], c["exponent"], verbose=False)
        assert y.dtype == torch.int64 and y.cpu().tolist() == c["y"].tolist(), (c["name"], c["num_groups"], c["outlier"])
    c = golden["bin"][1]
    y = du.bin_numeric_data(golden["columns"][c["x"]].tolist(), c["num_groups"], c["outlier"], c["exponent"])
    assert y == c["y"].tolist() and all(isinstance(v, int) for v in y)
    text = capsys.readouterr().out
    assert "Binning breakdown" in text and f"All {len(y)} elements successfully assigned to bins" in text
    e = golden["errors"]["all_filtered"]
    with pytest.raises(ValueError) as info:
        du.bin_numeric_data(golden["columns"][e["x"]].tolist(), e["num_groups"], e["outlier"], e["exponent"])
    assert str(info.value) == e["message"]


def test_bin_report_is_the_references(du, golden, capsys):
    """Same printed breakdown as data_utils.py:566-607 (checked against the oracle-independent count of groups)."""
    c = golden["bin"][0]
    x = golden["columns"][c["x"]].tolist()
    du.bin_numeric_data(x, c["num_groups"], c["outlier"], c["exponent"])
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.strip().startswith("Bin")]
    counts = {}
    for g in c["y"].tolist():
        counts[g] = counts.get(g, 0) + 1
    assert len(lines) == len(counts)
    for ln in lines:
        g = int(ln.split(":")[0].split()[-1])
        assert ln.rstrip().endswith(f"- {counts[g]} elements")


def test_vocabulary_golden(du, golden):
    for c in golden["vocab"]:
        x = dev(golden["columns"][c["x"]])
        tok, voc = du.numerical_representation(x)
        assert tok.dtype == torch.int32 and tok.cpu().tolist() == c["tokens"].tolist(), c["name"]
        assert voc == c["vocabulary"].tolist(), c["name"]
    tok, voc = du.numerical_representation([3, 1, 2, 3, 1])
    assert tok == [2, 0, 1, 2, 0] and voc == [1, 2, 3] and all(isinstance(v, int) for v in voc)
    assert du.numerical_representation([]) == ([], [])


def test_chain_golden_into_token_store(du, golden):
    from tradeaid_b200.feeder import TokenStore
    for c in golden["chain"]:
        x = dev(golden["columns"][c["x"]])
        n = x.numel()
        store = TokenStore([torch.zeros(n, dtype=torch.long), torch.zeros(n, dtype=torch.long)], "cuda")
        tok, voc = du.prepare_modality(x, c["params"], c["file_lengths"], tokens_out=store.store[1])
        assert tok.data_ptr() == store.store[1].data_ptr()
        assert store.store[1].cpu().tolist() == c["tokens"].tolist(), c["name"]
        assert [float(v) for v in voc] == c["vocabulary"].tolist(), c["name"]
        assert all(isinstance(v, int) for v in voc) == c["integer_vocabulary"], c["name"]
        assert int(store.store[0].abs().sum()) == 0


# ------------------------------------------------------------------------------------------------ against the oracle
def test_fresh_columns_against_oracle(du):
    from oracle import dataprep as oracle
    rng = random.Random(99)
    for trial in range(12):
        n = rng.choice([1, 2, 33, 257, 4000])
        col = [round(50 + rng.gauss(0, 20), rng.randint(0, 5)) or 1.0 for _ in range(n)]
        params = ["f", 1, None, trial % 2 == 0, rng.choice([None, 1, 2, 3]), rng.choice([0, 1, 2, 3]),
                  rng.choice([None, 4, 9]), None, False, "m", rng.choice([None, 2, 10]), rng.choice([None, 1.0, 2.5])]
        lens = None
        if params[3] and n >= 33:
            cut = rng.randint(1, n - 1)
            lens = [cut, n - cut]
        try:
            want_tok, want_voc = oracle.prepare_modality(col, params, lens)
        except ValueError as e:                     # e.g. every row outside the percentile band: same error on the device
            with pytest.raises(ValueError) as info:
                du.prepare_modality(col, params, lens)
            assert str(info.value) == str(e)
            continue
        tok, voc = du.prepare_modality(col, params, lens)
        assert tok.cpu().tolist() == want_tok, (trial, params)
        assert [float(v) for v in voc] == [float(v) for v in want_voc], (trial, params)   # (0.0 == -0.0: set() keeps either)


def test_order_statistics_against_sort(du):
    from tradeaid_b200 import lib
    torch.manual_seed(3)
    for n in [1, 2, 1000, 1 << 20]:
        x = torch.randn(n, dtype=torch.float64, device="cuda") * 1e3
        x[:: 7] = 0.0
        if n > 4:
            x[3] = -0.0
            x[1::11] = x[1]          # heavy duplicates
        srt = torch.sort(x).values.cpu()
        ranks = sorted({0, n - 1, n // 2, (n * 999) // 1000})
        st = torch.empty(lib.PREP_STATUS_WORDS, dtype=torch.int64, device="cuda")
        lib.prep_status_reset(st)
        out = torch.empty(len(ranks), dtype=torch.float64, device="cuda")
        ws = torch.empty(lib.prep_select_workspace_bytes(), dtype=torch.uint8, device="cuda")
        lib.prep_order_statistics(x, torch.tensor(ranks, device="cuda"), out, ws, st)
        assert out.cpu().tolist() == [float(srt[r]) for r in ranks], n
        assert st.tolist() == [lib.PREP_NO_ROW] * lib.PREP_STATUS_WORDS
        for q in [0, 0.1, 5, 50, 99.9, 100]:
            got = du.percentiles(x, [q])[0]
            assert got == float(np.percentile(x.cpu().numpy(), q)), (n, q)


# ------------------------------------------------------------------------------------------------ full size, by properties
def test_full_size_properties(du):
    """5 M rows (the c2 column is 1 M, c4's 50 M differs only in the grid-stride trip count): the vocabulary is strictly
    ascending, tokens index it back to the prepared column exactly, percent changes equal float64 torch arithmetic where
    no rounding tie is possible, ranged magnitudes stay inside [10^(d-1), 10^d) and are multiples of 10^-decimals."""
    n = 5_000_000
    g = torch.Generator(device="cuda").manual_seed(5)
    steps = torch.randn(n, dtype=torch.float64, device="cuda", generator=g) * 0.0004
    prices = (100.0 * torch.exp(torch.cumsum(steps, 0))).mul(100).round().div(100)
    pct = du.convert_to_percent_changes(prices, 2)
    raw = (prices[1:] - prices[:-1]) / prices[:-1] * 100
    assert float(pct[0]) == 0.0
    assert float((pct[1:] - raw).abs().max()) <= 0.005 + 1e-12
    scaled = pct * 100
    assert float((scaled - scaled.round()).abs().max()) < 1e-6            # two decimals
    ranged = du.range_numeric_data(prices, 2, 1)
    assert float(ranged.min()) >= 10.0 and float(ranged.max()) < 100.0
    assert float((ranged * 10 - (ranged * 10).round()).abs().max()) < 1e-9
    again = du.range_numeric_data(ranged, 2, 1)
    assert torch.equal(again, ranged)                                      # idempotent
    groups = du.bin_numeric_data(pct, 8, 0.1, 2.2, verbose=False, as_float=True)
    assert float(groups.min()) >= -8 and float(groups.max()) <= 8
    assert torch.equal(groups.sign(), pct.sign())                          # zero group only for zero, sign kept
    order = torch.argsort(pct)
    assert bool((groups[order][1:] >= groups[order][:-1]).all())          # monotone in the value
    for column, integer in [(ranged, False), (groups, True), (pct, False)]:
        tok, voc = du.numerical_representation(column, integer_vocabulary=integer)
        assert all(a < b for a, b in zip(voc, voc[1:]))
        v = torch.tensor(voc, dtype=torch.float64, device="cuda")
        assert torch.equal(v[tok.long()], column)
        assert int(tok.min()) == 0 and int(tok.max()) == len(voc) - 1


def test_vocabulary_table_growth(du):
    """More distinct values than half the first table: the insert reports it and the wrapper retries with a larger one."""
    n = 300_000
    x = torch.randperm(n, device="cuda").to(torch.float64) * 0.5 - 1000.0
    tok, voc = du.numerical_representation(x)
    assert len(voc) == n and voc[0] == -1000.0 and voc[-1] == (n - 1) * 0.5 - 1000.0
    assert torch.equal(tok.long(), ((x + 1000.0) * 2).long())


def test_prepared_dataset_feeds_get_batch(du):
    """Raw columns -> prepare_dataset -> training_utils.get_batch: the batches equal the ones drawn from the reference-style
    python lists built by the oracle chain (same seed), so prepared device rows are a drop-in for all_train_sets/val_sets."""
    from oracle import dataprep as oracle
    from tradeaid_b200 import training_utils as tu
    rng = random.Random(4)
    n, lens = 3000, [1000, 1200, 800]
    price = []
    p = 50.0
    for _ in range(n):
        p = max(1.0, p * (1 + rng.gauss(0, 0.01)))
        price.append(round(p, 2))
    volume = [float(rng.randint(1, 10 ** 5)) for _ in range(n)]
    params = [["p.csv", 1, None, True, None, 2, 8, None, False, "price pct", 0.5, 2.0],
              ["p.csv", 2, None, False, 2, 1, None, None, True, "volume", None, None]]
    train, val, vocabs = du.prepare_dataset([price, volume], params, lens, validation_size=0.1)
    want = [oracle.prepare_modality(c, q, lens) for c, q in zip([price, volume], params)]
    cut = int(n * (1 - 0.1))
    for i in range(2):
        assert train[i].is_cuda and train[i].dtype == torch.int32
        assert train[i].cpu().tolist() == want[i][0][:cut] and val[i].cpu().tolist() == want[i][0][cut:]
        assert [float(v) for v in vocabs[i]] == [float(v) for v in want[i][1]]

    def batches(train_sets, val_sets):
        tu.configure(device="cuda", batch_size=8, block_size=16, eval_iters=1)
        tu.all_train_sets, tu.all_val_sets = train_sets, val_sets
        tu.all_vocabularies = vocabs
        tu.all_modality_params = params
        tu.file_lengths = lens
        tu.num_modalities = 2
        tu.is_percents = True
        torch.manual_seed(123)
        out = []
        for split in ("train", "val", "train"):
            xb, yb = tu.get_batch(split, 0)
            out.append([t.cpu() for t in xb + yb])
        return out

    got = batches(train, val)
    ref = batches([w[0][:cut] for w in want], [torch.tensor(w[0][cut:], dtype=torch.long) for w in want])
    for a, b in zip(got, ref):
        for s, t in zip(a, b):
            assert torch.equal(s, t)


def test_range_with_decimal_places_none_on_the_device(du):
    """range_numeric_data(decimal_places=None) (data_utils.py:431: every element rounded to the fraction length of its own
    str()) on the device against the oracle: list form with Python ints and exponent-notation values, and a float64 column of
    100 k quoted prices / full-precision doubles, bit for bit."""
    from oracle import dataprep as O
    rng = random.Random(21)
    data = [12.5, 0.125, 100.0, 7, -3, 0.0, 1.5e-5, 2e16, -0.00042, 99.999, 1234.5678, 0.30000000000000004, 5, 1e-4]
    data += [round(rng.uniform(-3000, 3000), rng.randint(0, 5)) for _ in range(2000)]
    for nwd in (None, 2):
        got = du.range_numeric_data(list(data), nwd, None)
        assert same_bits(got, O.range_values(data, nwd, None)), nwd
    col = [round(rng.uniform(0.5, 900.0), rng.randint(1, 4)) for _ in range(60_000)]
    col += [rng.uniform(-1, 1) * 10.0 ** rng.randint(-3, 12) for _ in range(40_000)]
    x = dev(torch.tensor(col, dtype=torch.float64))
    for nwd in (None, 1, 3):
        got = du.range_numeric_data(x, nwd, None)
        assert same_bits(got.cpu().numpy(), O.range_values(col, nwd, None)), nwd
    # string modality through the chain: ids in the store row
    words = [rng.choice(["a", "b", "09:30", "Zed", ""]) for _ in range(1000)]
    tok, voc = du.prepare_modality(words, ["t.csv", 1, None, False, None, None, None, None, False, "label", None, None])
    assert voc == sorted(set(words)) and tok.is_cuda and tok.cpu().tolist() == [voc.index(w) for w in words]
