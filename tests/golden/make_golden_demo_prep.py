"""Generates tests/golden/demo_prep.pt: the reference's OWN demo (BASELINE.json configs[0]: examples/demo_input_schemas.yaml
on examples/demo_data/demo_stock.csv) through the reference's loader and processing functions, in main.py's order
(main.py:75-210, 278): raw column -> [percent changes at load time] -> ranging -> binning -> numerical_representation.
Stores the raw column (as the loader returns it without percent conversion), the legacy parameter lists, and the
reference's tokens / vocabularies — the demo's vocabulary sizes [57, 3] are the ones SURVEY.md quotes for c1.
Run in the build container only:  python tests/golden/make_golden_demo_prep.py"""
import contextlib
import io
import os
import sys

import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
os.chdir(REF)          # the schemas use paths relative to the repository root; nothing is written

import data_utils as ref  # noqa: E402
import file_cache  # noqa: E402

CSV = "./examples/demo_data/demo_stock.csv"
# schema.to_legacy_list layout (schema.py:213-248) of the two demo modalities
PARAMS = [
    [CSV, 13, True, False, 2, 1, None, None, True, "Demo Prices", None, None],
    [CSV, 13, True, True, None, 2, 3, None, False, "Demo Changes (%)", 0.1, None],
]


def main():
    out = {"params": [], "raw": [], "tokens": [], "vocabulary": [], "file_lengths": None}
    for p in PARAMS:
        with contextlib.redirect_stdout(io.StringIO()):
            raw, info = file_cache.load_file_data_cached([*p[:3], False, *p[4:10]])      # the column itself
            data, info = file_cache.load_file_data_cached(p[:10])                        # as main.py loads it
            if p[4] is not None or p[5] is not None:                                     # main.py:127 (decimals alone also "range")
                data = ref.range_numeric_data(data, p[4], p[5])
            if p[6] is not None:
                data = ref.bin_numeric_data(data, p[6], p[10] if p[10] is not None else 0.1, p[11] if p[11] is not None else 2.2)
            tok, voc = ref.numerical_representation(data)
        out["params"].append([None if k == 0 else v for k, v in enumerate(p)])           # the path stays behind
        out["raw"].append(torch.tensor([float(v) for v in raw], dtype=torch.float64))
        out["tokens"].append(torch.tensor(tok, dtype=torch.int64))
        out["vocabulary"].append(torch.tensor([float(v) for v in voc], dtype=torch.float64))
        out["file_lengths"] = [int(v) for v in info[1::2]]
    assert [len(v) for v in out["vocabulary"]] == [57, 3], [len(v) for v in out["vocabulary"]]
    torch.save(out, os.path.join(OUT, "demo_prep.pt"))
    print([len(v) for v in out["vocabulary"]], out["file_lengths"], os.path.getsize(os.path.join(OUT, "demo_prep.pt")))


if __name__ == "__main__":
    main()
