"""Row f-4 on the reference's own demo (BASELINE.json configs[0]): the two demo modalities of examples/demo_input_schemas.yaml
on examples/demo_data/demo_stock.csv, prepared on the device from the raw column, must give the tokens and vocabularies
the reference's loader + processing functions give (tests/golden/demo_prep.pt; vocabulary sizes [57, 3] = the c1 model)."""
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_reference_demo_columns_on_device():
    from tradeaid_b200 import data_utils as du
    g = load_golden("demo_prep.pt")
    cols = [raw.to("cuda") for raw in g["raw"]]
    train, val, vocabs = du.prepare_dataset(cols, g["params"], g["file_lengths"], validation_size=0.1)
    n = g["raw"][0].numel()
    cut = int(n * (1 - 0.1))
    for i in range(2):
        want = g["tokens"][i].tolist()
        assert train[i].cpu().tolist() == want[:cut] and val[i].cpu().tolist() == want[cut:]
        assert [float(v) for v in vocabs[i]] == g["vocabulary"][i].tolist()
    assert [len(v) for v in vocabs] == [57, 3]
