"""Offline differential soak (BUILD CONTAINER ONLY: imports /root/reference): random columns and random legacy parameter
lists through (1) the reference's data_utils functions in main.py's order, (2) the oracle restatement, (3) the product's
data_utils over the CPU emulation of the K9 kernels (the kernels' own element arithmetic built for the host). Tokens and
vocabularies must be identical, and so must the exceptions (all rows filtered, zero division).
usage: cd /tmp && python /root/repo/tests/soak/dataprep_chain_soak_vs_reference.py [trials] [seed]"""
import contextlib
import io
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, "/root/reference")
import data_utils as ref  # noqa: E402  (the reference's)
import emu_prep  # noqa: E402
from oracle import dataprep as oracle  # noqa: E402
from tradeaid_b200 import data_utils as du  # noqa: E402

du._ops_override = emu_prep
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 400
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)


def reference_chain(col, p, lens):
    x = list(col)
    with contextlib.redirect_stdout(io.StringIO()):
        if p[3] is True:
            out, at = [], 0
            for n in lens:
                out.extend(ref.convert_to_percent_changes(x[at:at + n], p[5] if p[5] else 2))
                at += n
            x = out
        if p[4] is not None or p[5] is not None:
            x = ref.range_numeric_data(x, p[4], p[5])
        if p[6] is not None:
            x = ref.bin_numeric_data(x, p[6], p[10] if p[10] is not None else 0.1, p[11] if p[11] is not None else 2.2)
        return ref.numerical_representation(x)


def column(n):
    kind = rng.choice(["walk", "uniform", "mixed", "ints", "small"])
    if kind == "walk":
        p, out = rng.uniform(1, 500), []
        for _ in range(n):
            p = max(0.01, p * (1 + rng.gauss(0, rng.choice([0.001, 0.01, 0.05]))))
            out.append(round(p, rng.choice([0, 1, 2, 4])))
        return out
    if kind == "uniform":
        return [rng.uniform(-100, 100) for _ in range(n)]
    if kind == "mixed":
        return [rng.choice([-1, 1]) * rng.uniform(0, 1) * 10 ** rng.randint(-6, 6) for _ in range(n)]
    if kind == "ints":
        return [float(rng.randint(-50, 50)) for _ in range(n)]
    return [round(rng.gauss(0, 0.01), 4) for _ in range(n)]


same = errors = 0
for trial in range(trials):
    n = rng.choice([1, 2, 7, 50, 300, 1200])
    col = column(n)
    lens = [n] if n < 7 or rng.random() < 0.5 else (lambda c: [c, n - c])(rng.randint(1, n - 1))
    p = ["f.csv", 1, None, rng.random() < 0.5, rng.choice([None, None, -1, 0, 1, 2, 3, 5]), rng.choice([None, 0, 1, 2, 3, 4, 6]),
         rng.choice([None, None, 1, 3, 6, 20]), None, False, "m", rng.choice([None, 0, 0.1, 5, 25, 50]), rng.choice([None, 1, 1.5, 2.2, 3.0])]
    if p[4] is not None and p[5] is None:
        p[5] = 2                                        # the str()-length form of range_numeric_data is outside the device path
    results = []
    for fn in (lambda: reference_chain(col, p, lens), lambda: oracle.prepare_modality(col, p, lens, on_zero="raise"),
               lambda: (lambda t, v: (t.tolist(), v))(*du.prepare_modality(col, p, lens, on_zero="raise"))):
        try:
            tok, voc = fn()
            # value equality (10 == 10.0, 0 == -0.0): range_numeric_data's clamps put Python ints into an otherwise float
            # vocabulary; the element TYPE is compared only after binning, where the reference's vocabulary is all ints
            results.append(("ok", [int(t) for t in tok], [float(v) for v in voc],
                            all(isinstance(v, int) for v in voc) if p[6] is not None else None))
        except (ValueError, ZeroDivisionError, OverflowError) as e:
            results.append(("err", type(e).__name__, str(e) if not isinstance(e, OverflowError) else ""))
    assert results[0] == results[1] == results[2], (trial, p, lens, col[:5], [r[:2] for r in results])
    same += results[0][0] == "ok"
    errors += results[0][0] == "err"
print(f"data preparation chain soak: {same} ok + {errors} error cases, reference = oracle = product host logic over the kernels' arithmetic")
