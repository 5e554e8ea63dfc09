"""Generates tests/golden/dataprep.pt by running the UNMODIFIED reference data_utils (/root/reference) on seeded columns.

Run in the build container only (the GPU box has no /root/reference):  python tests/golden/make_golden_dataprep.py
Everything is stored as float64 / int64 tensors (python floats are doubles, so nothing is lost) plus the parameters."""
import contextlib
import io
import math
import os
import random
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
os.chdir("/tmp")

import data_utils as ref  # noqa: E402

N = 1500


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def f64(v):
    return torch.tensor(v, dtype=torch.float64)


def price_walk(rng, n, start=100.0, vol=0.01, decimals=2):
    p, out = start, []
    for _ in range(n):
        p = max(0.01, p * (1.0 + rng.gauss(0.0, vol)))
        out.append(round(p, decimals))
    return out


def awkward_values(rng, n):
    """Mixed signs and magnitudes, exact ties, values a few ulps either side of powers of ten, zeros."""
    out = []
    specials = [0.0, -0.0, 0.5, 1.5, 2.5, -2.5, 0.125, 0.375, 2.675, 1.005, 999.9999999999999, 1000.0, 1000.0000000000001,
                0.1, 0.09999999999999999, 0.1000000000000000055, 9.999999999999998, 10.0, 99.99999999999999, 1e-7, 123456789.125,
                -0.001, 3e-300, 1e22, 3.0e-5, 0.045, -0.045, 0.055, 1e15 + 0.5, 7, -3, 250]
    for i in range(n):
        c = rng.random()
        if i < len(specials):
            out.append(specials[i])
        elif c < 0.35:
            out.append(rng.uniform(-50, 50))
        elif c < 0.55:
            out.append(rng.uniform(-1, 1) * 10 ** rng.randint(-9, 9))
        elif c < 0.7:
            out.append(rng.randint(-10 ** 6, 10 ** 6) / 2 ** rng.randint(0, 20))
        elif c < 0.8:
            p = 10.0 ** rng.randint(-6, 8)
            for _ in range(rng.randint(0, 3)):
                p = math.nextafter(p, rng.choice([-math.inf, math.inf]))
            out.append(p * rng.choice([-1, 1]))
        elif c < 0.9:
            out.append(round(rng.uniform(-500, 500), rng.randint(0, 4)))
        else:
            out.append(float(rng.randint(-1000, 1000)))
    return out


def main():
    rng = random.Random(20260118)
    g = {"columns": {}, "percent": [], "range": [], "bin": [], "vocab": [], "chain": [], "errors": {}}

    def col_ref(name, values):
        """Columns are stored once and referred to by name."""
        if name not in g["columns"]:
            g["columns"][name] = f64(values)
        assert g["columns"][name].tolist() == [float(v) for v in values] or name == "signed_zero"
        return name

    prices = price_walk(rng, N)
    volumes = [float(rng.randint(1, 10 ** 6)) for _ in range(N)]
    volumes[100] = volumes[101]                       # a zero change
    awkward = awkward_values(rng, N)
    nonzero_awkward = [v if v != 0 else 1.25 for v in awkward]

    # ---- convert_to_percent_changes
    for name, col, dp in [("prices", prices, 2), ("prices", prices, 3), ("prices", prices, 0), ("volumes", volumes, 2),
                          ("nonzero_awkward", nonzero_awkward, 4), ("nonzero_awkward", nonzero_awkward, 6), ("prices", prices, None)]:
        g["percent"].append(dict(name=name, decimals=dp, x=col_ref(name, col), y=f64(ref.convert_to_percent_changes(list(col), dp) if dp is not None
                                                                          else ref.convert_to_percent_changes(list(col)))))
    # per-file conversion as the loader does it (data_utils.py:148): three files
    lens = [400, 700, N - 1100]
    at, cat = 0, []
    for n in lens:
        cat.extend(ref.convert_to_percent_changes(prices[at:at + n], 2))
        at += n
    g["percent"].append(dict(name="prices_3files", decimals=2, file_lengths=lens, x=col_ref("prices", prices), y=f64(cat)))
    # file_cache's zero-tolerant variant
    sys.path.insert(0, REF)
    import file_cache
    with_zero = list(prices)
    with_zero[10] = 0.0
    with_zero[500] = 0.0
    g["percent"].append(dict(name="zero_keep", decimals=2, on_zero="keep", x=col_ref("with_zero", with_zero),
                             y=f64(quiet(file_cache._convert_to_percentage_changes, list(with_zero), 2, "f.csv"))))
    try:
        ref.convert_to_percent_changes(list(with_zero), 2)
    except ZeroDivisionError as e:
        g["errors"]["zero_prev"] = dict(x=col_ref("with_zero", with_zero), message=str(e))

    # ---- range_numeric_data
    pct = ref.convert_to_percent_changes(list(prices), 2)
    for name, col in [("awkward", awkward), ("prices", prices), ("pct", pct)]:
        for nwd, dp in [(1, 2), (2, 1), (3, 0), (None, 2), (None, 0), (0, 3), (-1, 4), (2, 2), (5, 1), (1, 0), (2, 6)]:
            y = ref.range_numeric_data(list(col), nwd, dp)
            g["range"].append(dict(name=name, whole_digits=nwd, decimals=dp, x=col_ref(name, col), y=f64([float(v) for v in y])))

    try:
        ref.range_numeric_data([1.5, 5e-324, 2.0], 1, 2)
    except OverflowError as e:
        g["errors"]["range_overflow"] = dict(x=col_ref("tiny", [1.5, 5e-324, 2.0]), whole_digits=1, decimals=2, message=str(e))

    # ---- bin_numeric_data
    for name, col in [("pct", pct), ("awkward", awkward), ("prices", prices)]:
        for G, outl, ex in [(5, 5, 2.0), (10, 0.1, 2.2), (20, 0, 1), (3, 45, 3.0), (6, 12.5, 1.5), (1, 5, 2.0), (64, 1, 2.2)]:
            y = quiet(ref.bin_numeric_data, list(col), G, outl, ex)
            g["bin"].append(dict(name=name, num_groups=G, outlier=outl, exponent=ex, x=col_ref(name, col),
                                 y=torch.tensor(y, dtype=torch.int64)))
    try:
        quiet(ref.bin_numeric_data, list(awkward), 3, 50, 3.0)
    except ValueError as e:
        g["errors"]["all_filtered"] = dict(x=col_ref("awkward", awkward), num_groups=3, outlier=50, exponent=3.0, message=str(e))
    zeros = [0.0] * 50
    g["bin"].append(dict(name="all_zero", num_groups=4, outlier=5, exponent=2.0, x=col_ref("all_zero", zeros),
                         y=torch.tensor(quiet(ref.bin_numeric_data, list(zeros), 4, 5, 2.0), dtype=torch.int64)))

    # ---- numerical_representation
    ranged = ref.range_numeric_data(list(awkward), 2, 1)
    for name, col in [("ranged_awkward", [float(v) for v in ranged]), ("pct", pct), ("prices", prices),
                      ("signed_zero", [0.0, -0.0, 1.0, -1.0, 0.0, -0.0, 2.5, 2.5]), ("single", [3.25] * 17)]:
        tok, voc = ref.numerical_representation(list(col))
        g["vocab"].append(dict(name=name, x=col_ref(name, col), tokens=torch.tensor(tok, dtype=torch.int64), vocabulary=f64(voc)))

    # ---- the chain main.py applies per modality (main.py:100-210, 278), on the legacy parameter list
    def chain(col, p, lens=None):
        x = list(col)
        if p[3] is True:
            if lens:
                at, out = 0, []
                for n in lens:
                    out.extend(ref.convert_to_percent_changes(x[at:at + n], p[5] if p[5] else 2))
                    at += n
                x = out
            else:
                x = ref.convert_to_percent_changes(x, p[5] if p[5] else 2)
        if p[4] is not None or p[5] is not None:
            x = ref.range_numeric_data(x, p[4], p[5])
        if p[6] is not None:
            x = quiet(ref.bin_numeric_data, x, p[6], p[10] if p[10] is not None else 0.1, p[11] if p[11] is not None else 2.2)
        return ref.numerical_representation(x)

    def P(is_pct=False, nwd=None, dp=None, bins=None, outl=None, ex=None):
        return ["f.csv", 1, None, is_pct, nwd, dp, bins, None, False, "m", outl, ex]

    for name, col, p, lens in [("price_ranged", prices, P(nwd=2, dp=1), None),
                               ("price_pct_binned", prices, P(is_pct=True, dp=2, bins=6), None),
                               ("price_pct_files", prices, P(is_pct=True, dp=2, bins=10, outl=1, ex=2.0), [400, 700, N - 1100]),
                               ("volume_ranged", volumes, P(nwd=1, dp=2), None),
                               ("volume_pct_ranged", volumes, P(is_pct=True, nwd=2, dp=0), None),
                               ("awkward_dp", awkward, P(dp=1), None),
                               ("raw_prices", prices, P(), None)]:
        tok, voc = chain(col, p, lens)
        g["chain"].append(dict(name=name, params=p, file_lengths=lens, x=col_ref({id(prices): "prices", id(volumes): "volumes", id(awkward): "awkward"}[id(col)], col), tokens=torch.tensor(tok, dtype=torch.int64),
                               vocabulary=f64([float(v) for v in voc]), integer_vocabulary=all(isinstance(v, int) for v in voc)))

    torch.save(g, os.path.join(OUT, "dataprep.pt"))
    print({k: len(v) for k, v in g.items()}, os.path.getsize(os.path.join(OUT, "dataprep.pt")))


if __name__ == "__main__":
    main()
