"""Pins oracle/restatement.py against fixtures produced by the reference itself (tests/golden/make_golden.py)."""
import pytest
import torch

from conftest import load_golden
from oracle import restatement as R

MODEL_CASES = ["demo", "m3", "m3_short", "hs32", "nocross"]


def _cfg(c):
    return R.Cfg(c["T"], c["C"], c["H"], c["L"], c["V"], c["cross"])


@pytest.mark.parametrize("name", MODEL_CASES)
def test_forward_matches_reference(name):
    g = load_golden(f"model_{name}.pt")
    cfg = _cfg(g["cfg"])
    logits, losses = R.forward(g["state_dict"], cfg, g["idx"], g["tgt"])
    for a, b in zip(logits, g["logits"]):
        assert torch.allclose(a, b, rtol=1e-5, atol=2e-6), (a - b).abs().max()
    for a, b in zip(losses, g["losses"]):
        assert abs(float(a) - float(b)) <= 1e-6 * max(1.0, abs(float(b)))
    # fp64 restatement: tighter agreement with the fp32 reference bounded by fp32 rounding only
    logits64, _ = R.forward(g["state_dict"], cfg, g["idx"], None, dtype=torch.float64)
    for a, b in zip(logits64, g["logits"]):
        assert (a.float() - b).abs().max() <= 5e-6


@pytest.mark.parametrize("name", MODEL_CASES)
def test_backward_matches_reference(name):
    g = load_golden(f"model_{name}.pt")
    cfg = _cfg(g["cfg"])
    _, _, grads = R.forward_backward(g["state_dict"], cfg, g["idx"], g["tgt"])
    assert set(grads) == set(g["grads"])
    for k, ref in g["grads"].items():
        if grads[k] is None:  # never-used parameter: the fixture stores zeros for a None grad
            assert ref.abs().max().item() == 0.0, k
            continue
        scale = ref.abs().max().item() + 1e-8
        assert (grads[k] - ref).abs().max().item() <= 2e-5 * scale + 1e-8, k


@pytest.mark.parametrize("name", MODEL_CASES)
def test_adamw_curve_matches_reference(name):
    g = load_golden(f"model_{name}.pt")
    cfg = _cfg(g["cfg"])
    params = {k: v.clone() for k, v in g["state_dict"].items()}
    m = {k: torch.zeros_like(v) for k, v in params.items()}
    v = {k: torch.zeros_like(p) for k, p in params.items()}
    curve = []
    for step in range(1, g["cfg"]["steps"] + 1):
        _, losses, grads = R.forward_backward(params, cfg, g["idx"], g["tgt"])
        curve.append(float(sum(losses)))
        R.adamw_step(params, grads, m, v, step, g["cfg"]["lr"])
    for a, b in zip(curve, g["curve"]):
        assert abs(a - b) <= 2e-5
    for k, ref in g["state_dict_after"].items():
        assert (params[k] - ref).abs().max().item() <= 2e-5, k


def test_state_dict_keys_of_random_init_match_reference():
    g = load_golden("model_m3.pt")
    cfg = _cfg(g["cfg"])
    sd = R.random_state_dict(cfg)
    ref_keys = [k for k in g["state_dict_keys"] if not k.endswith(".tril")]
    assert set(sd) == set(ref_keys)
    for k in ref_keys:
        assert sd[k].shape == g["state_dict"][k].shape, k


def test_indices_bit_exact():
    for c in load_golden("indices.pt"):
        torch.manual_seed(c["seed"])
        ix = R.batch_starting_indices(*c["args"])
        ix2 = R.batch_starting_indices(*c["args"])
        assert torch.equal(ix, c["ix"]) and torch.equal(ix2, c["ix2"]), c["args"]
        assert ix.dtype == torch.int64


def test_index_errors_match_reference_messages():
    with pytest.raises(ValueError, match="cannot be equal to or greater"):
        R.batch_starting_indices(8, 8, 4, "train", [8], False)
    with pytest.raises(ValueError, match="must be 'train' or 'val'"):
        R.batch_starting_indices(80, 8, 4, "test", [80], False)
    with pytest.raises(TypeError, match="'is_percents' must be a boolean"):
        R.batch_starting_indices(80, 8, 4, "train", [80], 0)
    with pytest.raises(ValueError, match="No valid starting positions"):
        R.batch_starting_indices(20, 8, 4, "train", [5, 5, 5, 5], False)


def test_get_batch_bit_exact():
    g = load_golden("get_batch.pt")
    full, n_val, T, B = g["full"], g["n_val"], g["block_size"], g["batch_size"]
    train = [d[:-n_val] for d in full]
    val = [d[-n_val:] for d in full]
    torch.manual_seed(g["seed"])
    for call in g["calls"]:
        data = train if call["split"] == "train" else val
        ix = R.batch_starting_indices(len(data[0]), T, B, call["split"], g["file_lengths"], False)
        xs, ys = R.gather_batch(data, ix, T)
        for a, b in zip(xs, call["xb"]):
            assert torch.equal(a, b)
        for a, b in zip(ys, call["yb"]):
            assert torch.equal(a, b)


def test_directional_metrics_match_reference():
    g = load_golden("metrics.pt")
    w, l, c, p = R.directional_metrics(g["logits"], g["xb"], g["yb"], g["vocabs"], g["params"])
    rw, rl, rc, rp = g["result"]
    assert w == rw and l == rl and p == rp
    for a, b in zip(c, rc):
        assert abs(a - b) <= 1e-5


def test_augmentation_matches_reference():
    """oracle add_rand_to_data_points vs outputs of data_utils.add_rand_to_data_points (tests/golden/augmentation.pt):
    values after one and three cumulative applications, the position of Python's `random` stream afterwards, errors."""
    import random
    g = load_golden("augmentation.pt")
    for c in g["cases"]:
        random.seed(c["seed"])
        work = list(c["data"])
        assert R.add_rand_to_data_points(work, c["size"], c["V"]) == c["after1"], c["name"]
        assert [random.getrandbits(32) for _ in range(4)] == c["tail1"]
        random.seed(c["seed"])
        work = list(c["data"])
        for _ in range(3):
            work = R.add_rand_to_data_points(work, c["size"], c["V"])
        assert work == c["after3"] and [random.getrandbits(32) for _ in range(4)] == c["tail3"]
    for size, exc, msg in g["errors"]:
        with pytest.raises({"TypeError": TypeError, "ValueError": ValueError}[exc]) as ei:
            R.add_rand_to_data_points([1, 2, 3], size, 10)
        assert str(ei.value) == msg


def test_loss_curve_prefix_matches_reference_training_run():
    """tests/golden/curve200.pt: the unmodified reference trained for 200 AdamW steps through its own get_batch. The
    oracle (index draw, gather, forward/backward, AdamW restatements chained) reproduces the first 40 losses to fp32
    round-off; the CUDA engine is held to the whole curve in tests/test_model_gpu.py."""
    from tradeaid_b200 import synthetic as syn
    g = load_golden("curve200.pt")
    c = g["cfg"]
    cfg = R.Cfg(c["T"], c["C"], c["H"], c["L"], c["V"], c["cross"])
    sd = R.random_state_dict(cfg, seed=0)
    for k in list(sd):
        if ".heads." in k and k.endswith(".2.weight") and "sa_layers" in k:
            sd[k] = sd[k] / 10
    data = [syn.token_stream("price", c["N"], 64, 1000), syn.token_stream("indicator", c["N"], 64, 1001)]
    m1 = {k: torch.zeros_like(v) for k, v in sd.items()}
    m2 = {k: torch.zeros_like(v) for k, v in sd.items()}
    torch.manual_seed(c["seed"])
    for step in range(1, 41):
        ix = R.batch_starting_indices(c["N"], c["T"], c["B"], "train", [c["N"]], False)
        xb, yb = R.gather_batch(data, ix, c["T"])
        _, losses, grads = R.forward_backward(sd, cfg, xb, yb)
        R.adamw_step(sd, grads, m1, m2, step, c["lr"])
        assert abs(float(sum(losses)) - g["curve"][step - 1]) <= 2e-4, (step, float(sum(losses)), g["curve"][step - 1])


@pytest.mark.parametrize("name", ["m3", "m3_short"])
def test_per_layer_activations_match_reference(name):
    """tests/golden/layer_activations.pt: forward hooks on the reference's blocks — the restatement must agree after
    every block, not only at the logits."""
    g = load_golden(f"model_{name}.pt")
    acts = load_golden("layer_activations.pt")[name]
    c = g["cfg"]
    cfg = R.Cfg(c["T"], c["C"], c["H"], c["L"], c["V"], c["cross"])
    _, per_layer = R.forward(g["state_dict"], cfg, g["idx"], None, return_hidden="layers")
    assert len(per_layer) == len(acts) == c["L"]
    for l in range(c["L"]):
        for a, b in zip(per_layer[l], acts[l]):
            assert (a - b).abs().max().item() <= 5e-6, (l, (a - b).abs().max().item())
